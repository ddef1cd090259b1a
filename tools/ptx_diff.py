#!/usr/bin/env python
"""Which kernels changed between two revisions?  Compiles every .cu of csrc/ to PTX at both revisions (nvcc -ptx,
compute_100a) and compares the kernels entry by entry after normalising symbol names, labels and line info.  PTX, not
SASS: ptxas is not deterministic for the tcgen05 kernels (two builds of the same source differ in instruction order and
register numbers), cicc is.

    python tools/ptx_diff.py <old-rev> [<new-rev, default: working tree>]  [--md out.md]
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = "self-similarity-grouping_b200/csrc"


def checkout(rev, dst):
    if rev is None:
        subprocess.check_call("cp -r %s/self-similarity-grouping_b200 %s/include %s/" % (ROOT, ROOT, dst), shell=True)
        os.makedirs(os.path.join(dst, "x"), exist_ok=True)
        return os.path.join(dst, "self-similarity-grouping_b200", "csrc")
    subprocess.check_call("git -C %s archive %s %s include | tar -x -C %s" % (ROOT, rev, CSRC, dst), shell=True)
    return os.path.join(dst, CSRC)


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def entries(csrc):
    out = {}
    for f in sorted(os.listdir(csrc)):
        if not f.endswith(".cu"):
            continue
        ptx = os.path.join(csrc, f + ".ptx")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-ptx", f, "-o", ptx],
                              cwd=csrc, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        s = open(ptx).read()
        for p in re.split(r"(?=\.visible \.entry )", s)[1:]:
            name = re.match(r"\.visible \.entry (\S+?)\(", p).group(1)
            body = p.replace(name, "K")
            body = re.sub(r"_ZZ\w+?E\d+(\w+?)\b", r"STATIC_\1", body)        # function-local statics (__shared__ arrays)
            body = re.sub(r"\$L__BB\d+_", "$L__BB_", body)
            body = re.sub(r"_Z\w+?_param_", "PARAM_", body)
            body = re.sub(r"__local_depot\d+", "__local_depot", body)
            body = "\n".join(l for l in body.splitlines() if not l.strip().startswith((".loc", "//", ".file")))
            out[name] = (f, body)
    return out


def key(dem):
    """demangled name -> name with the template arguments that were ADDED with a default stripped: compare by prefix."""
    return re.sub(r"\s+", "", dem)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    if md in args:
        args.remove(md)
    old_rev, new_rev = args[0], (args[1] if len(args) > 1 else None)
    with tempfile.TemporaryDirectory() as a, tempfile.TemporaryDirectory() as b:
        old, new = entries(checkout(old_rev, a)), entries(checkout(new_rev, b))
    dm = demangle(list(old) + list(new))
    old_by = {key(dm[n]): (n,) + v for n, v in old.items()}
    rows = []
    for n, (f, body) in sorted(new.items(), key=lambda kv: (kv[1][0], dm[kv[0]])):
        k = key(dm[n])
        match = old_by.get(k)
        if match is None:
            # a template that gained trailing parameters: the old name is a prefix of the new one up to the added
            # arguments, e.g. gemm_kernel<64, StagedEpi, true, false, 0> vs <..., 0, false>
            head = k.split("(")[0]
            cands = [ok for ok in old_by if ok.split("(")[0].rstrip(">") and head.startswith(ok.split("(")[0].rstrip(">"))
                     and head[len(ok.split("(")[0].rstrip(">")):] in (",false>", ",(bool)0>", ",0>")]
            match = old_by[cands[0]] if len(cands) == 1 else None
            if match is None:                  # same function name, different parameter list
                cands = [ok for ok in old_by if ok.split("(")[0] == head]
                match = old_by[cands[0]] if len(cands) == 1 else None
        short = dm[n].split("(")[0].replace("void ", "")
        if match is None:
            rows.append((f, short, "new"))
        else:
            rows.append((f, short, "identical" if match[2] == body else "CHANGED"))
    lines = ["# PTX of every kernel: %s -> %s\n" % (old_rev, new_rev or "working tree"),
             "`tools/ptx_diff.py` (nvcc -ptx, names / labels / line info normalised).  `identical` = the same PTX text, i.e. the "
             "same input to ptxas.\n",
             "| file | kernel | status |", "|---|---|---|"]
    for f, short, st in rows:
        lines.append("| %s | `%s` | %s |" % (f, short, st))
    n_id = sum(r[2] == "identical" for r in rows)
    lines.append("\n%d kernels: %d identical, %d changed, %d new." % (len(rows), n_id, sum(r[2] == "CHANGED" for r in rows),
                                                                   sum(r[2] == "new" for r in rows)))
    text = "\n".join(lines) + "\n"
    print(text)
    if md:
        with open(md, "w") as fo:
            fo.write(text)


if __name__ == "__main__":
    main()
