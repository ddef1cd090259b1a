#!/usr/bin/env python
"""Static evidence that the built library is Blackwell-native (no GPU needed).

Disassembles self-similarity-grouping_b200/lib/libssg_b200.so with `cuobjdump -sass` and counts, per kernel,
the SASS mnemonics B200_PROFILING.md names as proof: UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st),
UTMALDG/UTMASTG/UBLKCP (TMA), HMMA (legacy mma.sync: must be absent), plus the resource usage
(`cuobjdump -res-usage`: registers, static shared memory, spills via STL/LDL counts).

    python tools/sass_evidence.py > profiles/r01_sass_evidence.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "self-similarity-grouping_b200", "lib", "libssg_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP",
         "SYNCS", "HMMA", "DFMA", "DADD", "DMUL", "STL", "LDL", "REDUX", "SHFL", "VOTE")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name, width=96):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    cut = name.find("(")
    if cut > 0:
        name = name[:cut]
    return name if len(name) <= width else name[:width - 1] + "…"


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    arch = set()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    counts[cur][w] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(list(counts))
    print("# Static SASS evidence — `cuobjdump -sass` / `-res-usage` of `libssg_b200.so` (tools/sass_evidence.py)\n")
    print("Embedded architectures: %s.  %d kernels.  Mnemonics per B200_PROFILING.md: `UTC*MMA` = tcgen05.mma, "
          "`LDTM` = tcgen05.ld, `UTMALDG`/`UTMASTG` = TMA tensor load / store, `SYNCS` = mbarrier, `HMMA` = legacy "
          "mma.sync (must be 0), `STL`/`LDL` = local-memory (spill) traffic.\n" % (", ".join(sorted(arch)), len(counts)))
    tc = [k for k in counts if counts[k]["UTCHMMA"] or counts[k]["UTCQMMA"]]
    print("## tcgen05 / TMA kernels (%d)\n" % len(tc))
    print("| kernel | SASS instr | UTCHMMA | UTCBAR | LDTM | UTMALDG | UTMASTG | SYNCS | HMMA | STL+LDL | regs | static smem |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k in tc:
        c = counts[k]
        r = usage.get(k, ("?", "?"))
        print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %s | %s |" % (
            short(names[k]), c["_total"], c["UTCHMMA"] + c["UTCQMMA"], c["UTCBAR"], c["LDTM"], c["UTMALDG"],
            c["UTMASTG"], c["SYNCS"], c["HMMA"], c["STL"] + c["LDL"], r[0], r[1]))
    print("\n## all other kernels (%d)\n" % (len(counts) - len(tc)))
    print("| kernel | SASS instr | DFMA/DADD/DMUL | SHFL | VOTE | REDUX | HMMA | STL+LDL | regs | static smem |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for k in counts:
        if k in tc:
            continue
        c = counts[k]
        r = usage.get(k, ("?", "?"))
        print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %s | %s |" % (
            short(names[k]), c["_total"], c["DFMA"] + c["DADD"] + c["DMUL"], c["SHFL"], c["VOTE"], c["REDUX"],
            c["HMMA"], c["STL"] + c["LDL"], r[0], r[1]))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("\nLibrary totals: %d SASS instructions; UTCHMMA %d, LDTM %d, UTMALDG %d, UTMASTG %d, HMMA %d (legacy tensor path "
          "absent), local-memory instructions %d." % (tot["_total"], tot["UTCHMMA"] + tot["UTCQMMA"], tot["LDTM"],
                                                      tot["UTMALDG"], tot["UTMASTG"], tot["HMMA"], tot["STL"] + tot["LDL"]))


if __name__ == "__main__":
    main()
