#!/usr/bin/env python
"""Host-orchestration check of embed.cu without a GPU (TEST INFRASTRUCTURE).

embed.cu has no kernels of its own: it sequences the launchers of conv.cu over plan-owned buffers.  Here it is compiled
against tests/cpu_cuda/trace (every launcher logs its arguments, pointers as offsets into one arena), once from the
working tree and once from a git revision, and driven through the C ABI; the logs tell
  * whether a refactoring left the issued work unchanged (identical traces), and
  * -- by replaying a trace symbolically, one value id per image-pass and tensor -- whether a re-scheduled variant
    (SSG_L2_CHUNK) computes the same function of the same inputs without a buffer being overwritten before it is read.
"""
import ctypes
import hashlib
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "self-similarity-grouping_b200", "csrc")
OUT = os.path.join(HERE, "_build", "trace")


def build(tag, source_text):
    d = os.path.join(OUT, tag)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "embed.cpp"), "w") as f:
        f.write(source_text)
    for h in ("conv.h", "common.cuh"):
        with open(os.path.join(CSRC, h)) as f, open(os.path.join(d, h), "w") as g:
            g.write(f.read().replace('"../../include/ssg_b200.h"', '"ssg_b200.h"'))
    with open(os.path.join(HERE, "trace", "gemm_tc.cuh")) as f, open(os.path.join(d, "gemm_tc.cuh"), "w") as g:
        g.write(f.read())
    lib = os.path.join(d, "libembed_trace.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-D__CUDACC__", "-I" + d,
                    "-I" + os.path.join(HERE, "stub"), "-I" + os.path.join(ROOT, "include"), "-Wl,-Bsymbolic", "-o", lib,
                    os.path.join(d, "embed.cpp"), os.path.join(HERE, "trace", "trace_runtime.cpp")], check=True)
    return lib


def run(lib_path, n, batch=8, num_split=2, flip=1, env=None):
    """One process per run (the variants are read from the environment once per process)."""
    code = (
        "import ctypes, sys\n"
        "lib = ctypes.CDLL(%r)\n"
        "lib.trace_dump.restype = ctypes.c_char_p; lib.trace_alloc.restype = ctypes.c_void_p; lib.trace_alloc.argtypes = [ctypes.c_size_t]\n"
        "lib.ssg_last_error.restype = ctypes.c_char_p\n"
        "vp = ctypes.c_void_p\n"
        "lib.ssg_embed_plan_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]\n"
        "lib.ssg_embed_load_layer.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_float, vp]\n"
        "lib.ssg_embed_forward.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_size_t, ctypes.c_int, vp]\n"
        "plan = vp(); assert lib.ssg_embed_plan_create(ctypes.byref(plan), 0, %d, 256, 128) == 0\n"
        "dummy = lib.trace_alloc(1 << 20)\n"
        "for i in range(lib.ssg_embed_num_layers()):\n"
        "    assert lib.ssg_embed_load_layer(plan, i, dummy, dummy, dummy, dummy, dummy, 1e-5, None) == 0\n"
        "img = lib.trace_alloc(%d * 3 * 256 * 128 * 4); feat = lib.trace_alloc(3 * %d * 2048 * 4)\n"
        "lib.trace_reset()\n"
        "rc = lib.ssg_embed_forward(plan, img, %d, %d, 0, %d, feat, %d * 2048, 0, None)\n"
        "assert rc == 0, lib.ssg_last_error()\n"
        "sys.stdout.write(lib.trace_dump().decode())\n" % (lib_path, batch, n, n, n, num_split, flip, n))
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stderr
    return r.stdout.splitlines()


def _h(*parts):
    return hashlib.sha1(repr(parts).encode()).hexdigest()[:16]


def replay(lines, n, flip=1):
    """Symbolic execution: memory is a dict {arena offset of an image-pass slot: (value id, bytes)}.  Returns the value
    ids the pooled tail reads, one per image-pass."""
    mem = {}
    kv = lambda line: {k: int(v) for k, v in re.findall(r"(\w+)=(-?\d+)", line)}          # noqa: E731

    def rd(base, p, size):
        v = mem.get(base + p * size)
        return v[0] if v and v[1] == size else "GARBAGE@%d" % (base + p * size)

    def wr(base, p, size, val):
        # a write clobbers every slot it overlaps
        lo, hi = base + p * size, base + (p + 1) * size
        for a in [a for a, (_, s) in mem.items() if a < hi and a + s > lo]:
            del mem[a]
        mem[lo] = (val, size)
    out = None
    for line in lines:
        op, a = line.split(" ", 1)[0], kv(line)
        if op == "stem_prep":
            nb = a["n"] * (2 if a["flip"] else 1)
            for p in range(nb):
                wr(a["P"], p, 256 * 144 * 8, _h("prep", p % a["n"], p >= a["n"]))
        elif op == "stem64":
            for p in range(a["images"]):
                v = _h("stem", a["w"], a["b"], rd(a["P"], p, 256 * 144 * 8))
                if a["pool"] >= 0:
                    wr(a["pool"], p, 64 * 32 * 64 * 2, _h("pool", v))
                else:
                    wr(a["y"], p, 128 * 64 * 64 * 2, v)
        elif op == "conv1x1":
            # passes: the input slot at x tells the per-pass pixel count
            src = mem.get(a["x"])
            assert src is not None, "conv1x1 reads an unwritten buffer: " + line
            px = src[1] // (2 * a["cin"])
            nb = a["m"] // px
            for p in range(nb):
                ins = [rd(a["x"], p, px * a["cin"] * 2)]
                if a["res"] >= 0:
                    ins.append(rd(a["res"], p, px * a["cout"] * 2))
                wr(a["y"], p, px * a["cout"] * 2, _h("c1", a["w"], a["b"], a["relu"], a["cout"], *ins))
        elif op == "conv3x3":
            s = a["stride"]
            for p in range(a["B"]):
                v = rd(a["x"], p, a["H"] * s * a["W"] * s * a["cin"] * 2)
                wr(a["y"], p, a["H"] * a["W"] * a["cout"] * 2, _h("c3", a["w"], a["b"], a["relu"], s, a["cout"], v))
        elif op == "conv_fused_ds":
            s = a["stride"]
            for p in range(a["B"]):
                v1 = rd(a["t2"], p, a["H"] * a["W"] * a["mid"] * 2)
                v2 = rd(a["x"], p, a["H"] * s * a["W"] * s * a["cin"] * 2)
                wr(a["y"], p, a["H"] * a["W"] * a["cout"] * 2, _h("cds", a["w"], a["b"], s, a["cout"], v1, v2))
        elif op == "pooled_tail":
            nb = a["n"] * (2 if a["flip"] else 1)
            out = [rd(a["x"], p, 8 * 4 * 2048 * 2) for p in range(nb)]
        elif op in ("memcpy2d", "vec_add", "fold_bn", "fold_bn_stem"):
            pass
        else:
            raise AssertionError("unexpected launcher in the default configuration: " + line)
    assert out is not None, "no pooled_tail in the trace"
    return out


def head_source():
    with open(os.path.join(CSRC, "embed.cu")) as f:
        return f.read()


def git_source(rev):
    return subprocess.run(["git", "-C", ROOT, "show", rev + ":self-similarity-grouping_b200/csrc/embed.cu"], check=True,
                          capture_output=True, text=True).stdout


if __name__ == "__main__":
    new = build("head", head_source())
    t = run(new, 5)
    print(len(t), "launches;", "tail reads", replay(t, 5)[:2], "...")
