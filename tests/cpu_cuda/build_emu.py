#!/usr/bin/env python
"""Builds the CPU-emulated variant of the library's plain-CUDA translation units (TEST INFRASTRUCTURE).

api.cu, rerank.cu, cluster.cu, dist.cu, prof.cu and triplet.cu are copied with two mechanical rewrites --
    kernel<<<grid, block, smem, stream>>>(args)   ->  emu::launch([&]() { kernel(args); }, grid, block, smem, stream)
    extern __shared__ T name[];                   ->  T* name = (T*)emu::dyn_smem;
-- and compiled by g++ against tests/cpu_cuda/stub/cuda_runtime.h + emu.cpp (fibers, one host thread; see the stub's
header comment) into tests/cpu_cuda/_build/libssg_emu.so.  The tensor-core translation units (gemm_tc.cu, conv.cu,
embed.cu: tcgen05 / TMA, nothing to emulate) are left out, except the plain-CUDA operand split / centring kernels of
gemm_tc.cu; the approximate distance GEMM itself is a float stand-in (optionally perturbed within the certified error
bound), so that the tensor distance MODE -- candidate selection, exact re-scoring, certification, fallback -- runs too.  The plain-C harnesses of tests/c are then built against
it (same sources, `*_emu` binaries), which runs the real kernel source of the exact-mode re-ranking, eps and DBSCAN --
dense, row-sharded and sparse -- on the CPU.

    python tests/cpu_cuda/build_emu.py        # -> prints the paths of the binaries
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "self-similarity-grouping_b200", "csrc")
OUT = os.path.join(HERE, "_build")
UNITS = ["api.cu", "rerank.cu", "cluster.cu", "dist.cu", "prof.cu", "triplet.cu", "metrics.cu"]

STUBS = r'''// Stand-ins for the tensor-core distance GEMM (gemm_tc.cu: tcgen05 / TMA, not emulated).  The operand split, the
// centring and everything downstream of the GEMM (candidate selection, exact re-scoring, certification, fallback) are the
// library's real kernels; only the approximate d2 matrix itself is computed here, in plain float arithmetic on the
// same bf16x3 operands, optionally perturbed: SSG_EMU_GEMM_NOISE=f adds a deterministic pseudo-random error of up to
// f * E per entry, E = tensor_eps_rel(d) * (|x_i|^2 + max|y|^2) being the bound the certification assumes -- results
// must not depend on WHICH approximation within the bound the hardware produces.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"
namespace ssg {
static inline float bf(const void* p, size_t i) { return __bfloat162float(((const __nv_bfloat16*)p)[i]); }
int launch_gemm_dist(const void* a, const float* na, int m, const void* b, const float* nb, int n, int k, float* out,
                     size_t ldc, cudaStream_t, int sym) {
    const char* e = getenv("SSG_EMU_GEMM_NOISE");
    const double noise = e ? atof(e) : 0.0;
    const int d = k / 3;
    const double eps_rel = 1.18e-5 + 2.24e-8 * d;                 // api.cu tensor_eps_rel
    float nbmax = 0.f;
    for (int j = 0; j < n; ++j) nbmax = nb[j] > nbmax ? nb[j] : nbmax;
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) {
            if (sym && j < i) { out[(size_t)i * ldc + j] = out[(size_t)j * ldc + i]; continue; }
            float dot = 0.f;
            for (int q = 0; q < k; ++q) dot += bf(a, (size_t)i * k + q) * bf(b, (size_t)j * k + q);
            float v = fmaf(-2.0f, dot, na[i] + nb[j]);
            if (noise != 0.0) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                const double u = (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
                v += (float)(u * noise * eps_rel * ((double)na[i] + (double)nbmax));
            }
            out[(size_t)i * ldc + j] = v;
        }
    return SSG_OK;
}
int launch_sqdist_tensor(const float*, int, const float*, int, int, float*, size_t, cudaStream_t) {
    return ssg_set_error(SSG_ERR_UNSUPPORTED, "CPU emulation: ssg_sqdist(mode = TENSOR) is not emulated");
}
}  // namespace ssg
'''


def _match_forward(s, i, open_ch, close_ch):
    """s[i] == open_ch -> index just past the matching close_ch."""
    depth = 0
    while True:
        c = s[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def rewrite(src):
    src = re.sub(r"extern\s+__shared__\s+([A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2 = (\1*)emu::dyn_smem;", src)
    out, pos = [], 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            out.append(src[pos:])
            return "".join(out)
        # kernel expression: identifier (with ::) optionally followed by a template argument list, just before <<<
        j = k
        while src[j - 1].isspace():
            j -= 1
        if src[j - 1] == ">":                          # template arguments: walk back to the matching <
            depth, j = 0, j - 1
            while True:
                if src[j] == ">":
                    depth += 1
                elif src[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                j -= 1
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] in "_:"):
            j -= 1
        kern = src[j:k].strip()
        e = src.index(">>>", k)
        cfg = src[k + 3:e]
        a = e + 3
        while src[a].isspace():
            a += 1
        assert src[a] == "(", "launch without an argument list near: " + src[k - 40:k + 40]
        b = _match_forward(src, a, "(", ")")
        args = src[a + 1:b - 1]
        out.append(src[pos:j])
        out.append("emu::launch([&]() { %s(%s); }, %s)" % (kern, args, cfg))
        pos = b


def _up_to_date(outputs):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "emu.cpp"), os.path.abspath(__file__),
            os.path.join(HERE, "stub", "cuda_runtime.h"), os.path.join(HERE, "jaccard_big.cpp"),
            os.path.join(ROOT, "include", "ssg_b200.h"), os.path.join(ROOT, "tests", "c", "shard_check.c"),
            os.path.join(ROOT, "tests", "c", "sparse_check.c")]
    if not all(os.path.isfile(o) for o in outputs):
        return False
    return min(os.path.getmtime(o) for o in outputs) > max(os.path.getmtime(d) for d in deps)


PLAIN_FAULTS = {
    # name -> (translation unit, text, replacement): deliberate races, to show that running the kernels under different
    # thread schedules (emu.cpp: SSG_EMU_SCHED / ssg_emu_set_sched) notices them (tests/test_cpu_emulated_kernels.py)
    # thread 0 publishes the selection state, everybody reads it: without the barrier only "thread 0 first" works
    "eps_pick_no_barrier": ("cluster.cu",
                            "        if (n_pairs >= 0 && top > total) { state[7] = 1ull; state[1] = 0ull; }\n    }\n    __syncthreads();\n",
                            "        if (n_pairs >= 0 && top > total) { state[7] = 1ull; state[1] = 0ull; }\n    }\n"),
}


def build(verbose=False, force=False, fault=None):
    """fault: a key of PLAIN_FAULTS -> only the library, as _build/libssg_emu_<fault>.so"""
    tag = "_" + fault if fault else ""
    outputs = [os.path.join(OUT, "libssg_emu%s.so" % tag)] + ([] if fault else [os.path.join(OUT, b) for b in
                                                      ("shard_check_emu", "sparse_check_emu", "jaccard_big")])
    if not force and _up_to_date(outputs):
        return outputs[0], outputs[1:]
    os.makedirs(os.path.join(OUT, "src" + tag), exist_ok=True)
    srcs = []
    for u in UNITS:
        with open(os.path.join(CSRC, u)) as f:
            text = f.read()
        if fault and PLAIN_FAULTS[fault][0] == u:
            assert text.count(PLAIN_FAULTS[fault][1]) == 1, "fault anchor not found: " + fault
            text = text.replace(PLAIN_FAULTS[fault][1], PLAIN_FAULTS[fault][2])
        text = rewrite(text)
        assert "<<<" not in text
        dst = os.path.join(OUT, "src" + tag, u.replace(".cu", "_emu.cpp"))
        with open(dst, "w") as f:
            f.write(text)
        srcs.append(dst)
    stub = os.path.join(OUT, "src" + tag, "tc_stubs_emu.cpp")
    with open(stub, "w") as f:
        f.write(STUBS)
    # the plain-CUDA front end of the tensor distance mode (column mean, bf16x3 operand split) lives in gemm_tc.cu next
    # to the tcgen05 code: take just that region
    with open(os.path.join(CSRC, "gemm_tc.cu")) as f:
        g = f.read()
    region = g[g.index("constexpr int MEAN_GROUPS"):g.index("struct EpiDist {")]
    prep = os.path.join(OUT, "src" + tag, "dist_prep_emu.cpp")
    with open(prep, "w") as f:
        f.write('#include <cuda_bf16.h>\n#include "common.cuh"\n#include "kernels.h"\nnamespace ssg {\n'
                + rewrite(region) + "\n}  // namespace ssg\n")
    srcs += [stub, prep, os.path.join(HERE, "emu.cpp")]
    lib = outputs[0]
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing", "-D__CUDACC__", "-w",
             "-I" + os.path.join(HERE, "stub"), "-I" + CSRC, "-I" + os.path.join(ROOT, "include")]
    # -Bsymbolic: the library's cudaMalloc / cudaSetDevice / ... must bind to ITS stand-ins even inside a process that has
    # the real libcudart loaded (pytest imports torch)
    cmd = ["g++"] + flags + ["-shared", "-Wl,-Bsymbolic", "-o", lib] + srcs
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    bins = []
    if fault:
        return lib, bins
    for h in ("shard_check", "sparse_check"):
        exe = os.path.join(OUT, h + "_emu")
        subprocess.run(["gcc", "-std=c99", "-O1", "-I" + os.path.join(HERE, "stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", h + ".c"), "-o", exe, "-L" + OUT, "-lssg_emu", "-lm",
                        "-Wl,-rpath," + OUT], check=True)
        bins.append(exe)
    # C++ driver on the library's internal launch wrappers: sparse vs dense Jaccard at n > 8192 (slow: opt-in test)
    exe = os.path.join(OUT, "jaccard_big")
    subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I" + os.path.join(HERE, "stub"), "-I" + CSRC,
                    "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "jaccard_big.cpp"), "-o", exe, "-L" + OUT,
                    "-lssg_emu", "-Wl,-rpath," + OUT], check=True)
    bins.append(exe)
    return lib, bins


FAULTS = {
    # name -> (text in gemm_tc.cuh, replacement): deliberate protocol violations, to show that the late completion model
    # of emu_tc.cpp notices them (tests/test_cpu_emulated_tensor_kernels.py)
    "epi2_no_store_wait": ("if (leader) tma_store_wait_read<0>();             // store g-1 has left buffer (g+1) & 1", ""),
    "stage_freed_early": ("umma_commit(&empty_bar[stage]);                        // smem slot free once the MMAs retire",
                          "mbar_arrive(&empty_bar[stage]);"),
}


def build_tc(verbose=False, force=False, fault=None):
    """The whole library, tensor-core kernels included, against the FUNCTIONAL emulation of tcgen05 / TMA / mbarrier
    (stub_tc/tc_common.cuh): gemm_tc.cuh compiles unchanged except for the five helpers it defines with inline PTX
    (TMA store, named barrier), which are cut from its text.  -> _build/libssg_emu_tc.so (same C ABI, every entry point)."""
    out = os.path.join(OUT, "libssg_emu_tc%s.so" % ("_" + fault if fault else ""))
    src_dir = os.path.join(OUT, "src_tc" + ("_" + fault if fault else ""))
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.abspath(__file__), os.path.join(HERE, "emu.cpp"),
            os.path.join(HERE, "emu_tc.cpp")] + [os.path.join(HERE, d, f) for d in ("stub", "stub_tc")
                                                   for f in os.listdir(os.path.join(HERE, d))]
    if not force and os.path.isfile(out) and os.path.getmtime(out) > max(os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(src_dir, exist_ok=True)
    srcs = []
    for u in [f for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]:
        with open(os.path.join(CSRC, u)) as f:
            text = rewrite(f.read())
        dst = os.path.join(src_dir, u.replace(".cu", "_emu.cpp"))
        with open(dst, "w") as f:
            f.write(text)
        srcs.append(dst)
    with open(os.path.join(CSRC, "gemm_tc.cuh")) as f:
        h = f.read()
    a = h.index("__device__ __forceinline__ void tma_store_2d(")
    b = h.index("template <int BN, class Epi, bool STAGED, bool KHS = false, int VAR = VAR_NONE, bool EPI2 = false>\n__global__")
    h = (h[:a] + "// (TMA store helpers: stub_tc/tc_common.cuh)\n"
         "__device__ __forceinline__ void epi_bar_sync() { emu::named_barrier(1, 32 * EPI_WARPS); }\n\n" + h[b:])
    assert "asm volatile" not in h
    if fault:
        good, bad = FAULTS[fault]
        assert h.count(good) == 1, "fault anchor not found: " + fault
        h = h.replace(good, bad)
    with open(os.path.join(src_dir, "gemm_tc.cuh"), "w") as f:
        f.write(rewrite(h))
    # the two-CTA (cta_group::2) kernel needs thread-block clusters, which the emulator does not model: its launcher is
    # replaced by a stand-in that reports SSG_ERR_UNSUPPORTED (the variant is opt-in on the GPU as well)
    with open(os.path.join(src_dir, "gemm_tc2.cuh"), "w") as f:
        f.write('#pragma once\n#define SSG_PAIR_KERNEL_UNAVAILABLE 1\n#include "gemm_tc.cuh"\nnamespace ssg { namespace tc {\n'
                'template <int BN, bool RES, bool KHS = false>\n'
                'int launch_gemm2_op(const AOperand&, int, const void*, int, int, const StagedEpi&, cudaStream_t) {\n'
                '    return ssg_set_error(SSG_ERR_UNSUPPORTED, "the two-CTA kernel is not available under emulation");\n}\n'
                '} }\n')
    for hname, hdir in (("tc_common.cuh", os.path.join(HERE, "stub_tc")), ("conv.h", CSRC), ("kernels.h", CSRC)):
        with open(os.path.join(hdir, hname)) as f, open(os.path.join(src_dir, hname), "w") as g:
            g.write(f.read())
    with open(os.path.join(CSRC, "common.cuh")) as f, open(os.path.join(src_dir, "common.cuh"), "w") as g:
        g.write(f.read().replace('"../../include/ssg_b200.h"', '"ssg_b200.h"'))
    flags = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-fno-strict-aliasing", "-D__CUDACC__", "-w",
             "-I" + src_dir, "-I" + os.path.join(HERE, "stub_tc"), "-I" + os.path.join(HERE, "stub"),
             "-I" + os.path.join(ROOT, "include")]
    cmd = ["g++"] + flags + ["-shared", "-Wl,-Bsymbolic", "-o", out] + srcs + [os.path.join(HERE, "emu.cpp"),
                                                                             os.path.join(HERE, "emu_tc.cpp")]
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return out


if __name__ == "__main__":
    lib, bins = build(verbose=True)
    print(lib)
    for b in bins:
        print(b)
    print(build_tc(verbose=True))
