"""CPU: the tcgen05 / TMA kernels EXECUTED on the host.  tests/cpu_cuda/stub_tc/tc_common.cuh is a functional stand-in
for the Blackwell primitives csrc/tc_common.cuh wraps in inline PTX -- mbarriers (arrival counts, transaction bytes,
phases), TMA tensor loads and stores (boxes, traversal strides, out-of-bounds zero fill, 64- / 128-byte swizzle),
tcgen05.mma on shared-memory matrix descriptors into an emulated TMEM, tcgen05.commit / ld, named barriers -- so that
gemm_tc.cuh (the warp-specialised persistent GEMM with all its variants), conv.cu and embed.cu compile and run
unchanged on cooperative fibers.  What it establishes: operand staging, descriptor arithmetic, pipeline protocol (a
missed arrival is a detected deadlock), epilogues and the host orchestration compute the right function; what it cannot:
timing, data races, the tensor core's internal accumulation order (results are compared with float references to a
tolerance, and BETWEEN variants bit for bit).

The asynchronous units have two completion models (tests/cpu_cuda/emu_tc.cpp): "eager" -- a TMA load / store or an MMA
takes effect when it is issued -- and "late" -- it takes effect at the last moment the protocol allows (when the
mbarrier it signals is polled, when cp.async.bulk.wait_group stops tolerating it) -- plus "mixed", a seeded draw between
the two per operation.  A correct kernel computes the same
bytes under both; the fault-injection test at the end shows that the late model catches a stage released before its
MMAs retired and a staging buffer rewritten under an in-flight TMA store, which the eager model cannot see."""
import ctypes
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_cuda"))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


def to_bf16(a):
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x7fff + ((u >> 16) & 1)) >> 16).astype(np.uint16)


def from_bf16(h):
    return (h.astype(np.uint32) << 16).view(np.float32)


def conv_ref(x, w, bias, k, stride, res, relu):
    B, H, W, C = x.shape
    co, OH, OW, pad = w.shape[0], H // stride, W // stride, k // 2
    xp = np.zeros((B, H + 2 * pad, W + 2 * pad, C), np.float64)
    xp[:, pad:pad + H, pad:pad + W] = x
    y = np.zeros((B, OH, OW, co), np.float64)
    for kh in range(k):
        for kw in range(k):
            y += xp[:, kh:kh + H:stride, kw:kw + W:stride][:, :OH, :OW] @ w[:, kh, kw, :].T.astype(np.float64)
    y += bias
    if res is not None:
        y += res
    return np.maximum(y, 0) if relu else y


@pytest.fixture(scope="module")
def tc_lib():
    import build_emu
    from ssg_b200 import _lib as L
    lib = ctypes.CDLL(build_emu.build_tc())
    for name, (res, args) in L.PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


@pytest.mark.parametrize("late", [0, 1], ids=["eager", "late"])
@pytest.mark.parametrize("B,H,W,cin,cout,k,stride,use_res,relu", [
    (2, 8, 16, 64, 64, 1, 1, True, True),        # 128x64 tiles, whole-tile residual double buffer
    (2, 8, 16, 64, 128, 1, 1, True, False),      # 128x128 tiles
    (3, 8, 16, 256, 256, 1, 1, True, True),      # 128x256 tiles, residual sub-tile ring
    (3, 8, 16, 256, 512, 1, 1, False, True),     # 128x256 tiles, no residual
    (1, 16, 16, 128, 128, 3, 1, False, True),    # implicit 3x3: shifted 4-D TMA boxes, halo zero fill
    (1, 8, 32, 64, 64, 3, 1, False, True),       # kernel-row sharing (one haloed box serves three kernel rows)
    (2, 16, 32, 128, 128, 3, 2, False, True),    # stride 2 through element-strided TMA boxes
    (2, 16, 16, 256, 512, 1, 2, False, False),   # 1x1 stride 2 (downsample branch)
])
def test_convolution_kernels_against_float_reference(tc_lib, B, H, W, cin, cout, k, stride, use_res, relu, late):
    tc_lib.ssg_emu_set_async(late)
    rng = np.random.RandomState(B * 1000 + cin + cout + k)
    x = from_bf16(to_bf16(rng.randn(B, H, W, cin)))
    w = from_bf16(to_bf16(rng.randn(cout, k, k, cin) / np.sqrt(cin * k * k)))
    bias = (rng.randn(cout) * 0.1).astype(np.float32)
    OH, OW = H // stride, W // stride
    res = from_bf16(to_bf16(rng.randn(B, OH, OW, cout))) if use_res else None
    xb, wb = to_bf16(x), to_bf16(w)
    rb = to_bf16(res) if use_res else None
    y = np.zeros((B, OH, OW, cout), np.uint16)
    scratch = np.zeros(x.size + 64, np.uint16)
    rc = tc_lib.ssg_op_conv(xb.ctypes.data, B, H, W, cin, k, stride, wb.ctypes.data, bias.ctypes.data, cout,
                            rb.ctypes.data if use_res else None, int(relu), y.ctypes.data, scratch.ctypes.data, None)
    tc_lib.ssg_emu_set_async(0)
    assert rc == 0, tc_lib.ssg_last_error().decode()
    want = conv_ref(x, w, bias, k, stride, res, relu)
    assert np.abs(from_bf16(y) - want).max() <= 0.01 * np.abs(want).max() + 0.02      # bf16 output rounding


@pytest.mark.parametrize("B,H,W,cin,cout,k,stride,use_res", [
    (3, 8, 16, 256, 256, 1, 1, True),            # residual sub-tile ring, 4 K blocks over 3 stages
    (1, 8, 32, 64, 64, 3, 1, False),             # kernel-row sharing
    (2, 16, 32, 128, 128, 3, 2, False),          # stride-2 boxes, 18 K blocks
])
def test_convolution_kernels_are_byte_identical_under_every_schedule_and_completion_model(tc_lib, B, H, W, cin, cout, k,
                                                                                        stride, use_res):
    """The warp-specialised GEMM under the thread schedules of tests/cpu_cuda/emu.cpp (who of producer, MMA issuer and the
    eight epilogue warps runs first after every yield) crossed with the two completion models: same bytes.  The random
    schedule is warp-granular here: the lanes of a warp poll an mbarrier as one instruction on the hardware, and the
    "all lanes wait, lane 0 issues and commits, __syncwarp" idiom of the MMA warp depends on it -- a lane-granular random
    order, which lets the producer refill a stage between lane 0's poll and lane 1's, deadlocks by construction (the
    emulator reports it)."""
    rng = np.random.RandomState(11)
    xb = to_bf16(rng.randn(B, H, W, cin))
    wb = to_bf16(rng.randn(cout, k, k, cin) / np.sqrt(cin * k * k))
    bias = (rng.randn(cout) * 0.1).astype(np.float32)
    OH, OW = H // stride, W // stride
    rb = to_bf16(rng.randn(B, OH, OW, cout)) if use_res else None
    outs = {}
    try:
        for sched, (mode, seed) in {"forward": (0, 0), "reverse": (1, 0), "random warps": (3, 5)}.items():
            for late in (0, 1, 2):                           # eager, late, mixed (seeded draw per operation)
                tc_lib.ssg_emu_set_sched(mode, seed)
                tc_lib.ssg_emu_set_async(late)
                tc_lib.ssg_emu_seed_async(7 + seed)
                y = np.zeros((B, OH, OW, cout), np.uint16)
                scratch = np.zeros(xb.size + 64, np.uint16)
                rc = tc_lib.ssg_op_conv(xb.ctypes.data, B, H, W, cin, k, stride, wb.ctypes.data, bias.ctypes.data, cout,
                                        rb.ctypes.data if use_res else None, 1, y.ctypes.data, scratch.ctypes.data, None)
                assert rc == 0, tc_lib.ssg_last_error().decode()
                outs[(sched, late)] = y
    finally:
        tc_lib.ssg_emu_set_sched(0, 0)
        tc_lib.ssg_emu_set_async(0)
    assert outs[("forward", 0)].any()
    for key, y in outs.items():
        assert np.array_equal(y, outs[("forward", 0)]), key


def _embed(tmp_path, name, env, n=1):
    out = str(tmp_path / (name + ".npy"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cpu_cuda", "run_embed_emu.py"), str(n), out],
                       capture_output=True, text=True, env=dict(os.environ, **env), timeout=1200)
    assert r.returncode == 0, r.stdout + r.stderr
    rel = float([l for l in r.stdout.splitlines() if l.startswith("rel err")][0].split(":")[1].split()[0])
    return np.load(out), rel


def test_whole_trunk_against_reference_golden_and_variants_bit_identical(tmp_path):
    """The full ResNet-50 forward of ssg_embed_forward (stem with resident weights, parity-plane staging and fused
    max-pool; kernel-row sharing; element-strided stride-2 boxes; K-concatenated downsample; residual ring; pooled tail)
    on the reference's golden image and weights: features within the GPU smoke tolerance of the reference's, and the
    opt-in variants -- one-barrier epilogue, L2-resident chunking (direct and through graph capture), plain stem --
    bit-identical to the default.  The variants run under the LATE completion model (module docstring), the default under
    the eager one: the same bytes under both is the protocol check."""
    import build_emu
    from concurrent.futures import ThreadPoolExecutor
    build_emu.build_tc()                                     # once, before the runs start side by side
    variants = (("default_late", {}),
                ("default_reverse_order", {"SSG_EMU_SCHED": "reverse"}),          # thread / block schedules of emu.cpp
                ("default_random_warps", {"SSG_EMU_SCHED": "warps:3"}),
                ("default_mixed_completion", {"SSG_EMU_ASYNC": "mixed:3", "SSG_EMU_SCHED": "warps:4"}),
                ("epi2_chunk", {"SSG_CONV_EPI2": "1", "SSG_L2_CHUNK": "1", "SSG_L2_GRAPH": "0"}),
                ("chunk_graph", {"SSG_L2_CHUNK": "1", "SSG_L2_GRAPH": "1"}),
                ("plain_stem", {"SSG_STEM_BRES": "0", "SSG_STEM_POOL": "0", "SSG_CONV_BN256_RES": "0"}),
                ("khs_streamed_weights", {"SSG_KHS_BRES": "0"}),            # default: VAR_KHSB (weights resident)
                ("no_chain", {"SSG_CONV_CHAIN": "0"}),                      # default: layer-1 conv3 + next conv1 chained
                ("no_chain_reverse", {"SSG_CONV_CHAIN": "0", "SSG_EMU_SCHED": "reverse"}))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:      # one subprocess each
        first = pool.submit(_embed, tmp_path, "default", {})
        rest = [(name, pool.submit(_embed, tmp_path, name, dict({"SSG_EMU_ASYNC": "late"}, **env))) for name, env in variants]
        base, rel = first.result()
        assert rel < 3e-2 and np.isfinite(base).all()
        for name, fut in rest:
            assert np.array_equal(fut.result()[0], base), name


def test_distance_gemm_kernel_and_its_symmetric_variant(tmp_path):
    """The real distance GEMM kernel (EpiDist / EpiDistSym epilogues) under emulation inside the tensor distance mode:
    outputs bit-equal to the exact mode, with and without SSG_DIST_SYM (read once per process -> subprocesses), under
    eager completion in thread order and under late completion with randomly ordered warps."""
    script = (
        "import sys, os, ctypes, numpy as np\n"
        "sys.path[:0] = %r\n"
        "import build_emu\n"
        "from ssg_b200 import _lib as L\n"
        "from oracle import ssg_oracle as O\n"
        "lib = ctypes.CDLL(build_emu.build_tc())\n"
        "for nm in ('ssg_rerank_plan_create', 'ssg_rerank_plan_destroy', 'ssg_rerank_run', 'ssg_rerank_get_stage', 'ssg_last_error'):\n"
        "    getattr(lib, nm).restype, getattr(lib, nm).argtypes = L.PROTOTYPES[nm]\n"
        "n, ns, d = 300, 130, 64\n"
        "t, _ = O.synth_features(n, d, 3, per_cluster=12); s, _ = O.synth_features(ns, d, 4, noise=0.6)\n"
        "out = {}\n"
        "for mode in (0, 1):\n"
        "    plan = ctypes.c_void_p(); assert lib.ssg_rerank_plan_create(ctypes.byref(plan), 0, n, ns, d) == 0\n"
        "    f = np.empty((n, n))\n"
        "    rc = lib.ssg_rerank_run(plan, s.ctypes.data, ns, t.ctypes.data, n, d, 20, 6, 0.1, mode, f.ctypes.data, None, None)\n"
        "    assert rc == 0, lib.ssg_last_error().decode()\n"
        "    out[mode] = f\n"
        "    lib.ssg_rerank_plan_destroy(plan)\n"
        "print('SAME' if np.array_equal(out[0], out[1]) else 'DIFFERENT')\n"
        % ([ROOT, os.path.join(ROOT, "self-similarity-grouping_b200"), os.path.join(ROOT, "tests", "cpu_cuda")],))
    import build_emu
    from concurrent.futures import ThreadPoolExecutor
    build_emu.build_tc()
    cases = (("0", "eager"), ("1", "eager"), ("0", "late"), ("1", "late"))
    with ThreadPoolExecutor(max_workers=min(4, os.cpu_count() or 1)) as pool:
        runs = [pool.submit(subprocess.run, [sys.executable, "-c", script], capture_output=True, text=True, timeout=1200,
                            env=dict(os.environ, SSG_DIST_SYM=sym, SSG_EMU_ASYNC=model,
                                     SSG_EMU_SCHED="warps:2" if model == "late" else "forward")) for sym, model in cases]
        for (sym, model), fut in zip(cases, runs):
            r = fut.result()
            assert r.returncode == 0 and "SAME" in r.stdout, (sym, model, r.stdout + r.stderr)


FAULT_SCRIPT = r"""
import sys, os, ctypes, numpy as np
sys.path[:0] = %r
import build_emu
from ssg_b200 import _lib as L
from test_cpu_emulated_tensor_kernels import to_bf16, from_bf16, conv_ref
lib = ctypes.CDLL(build_emu.build_tc(fault=sys.argv[1] if sys.argv[1] != "none" else None))
lib.ssg_op_conv.restype, lib.ssg_op_conv.argtypes = L.PROTOTYPES["ssg_op_conv"]
B, H, W, cin, cout = 9, 8, 16, 512, 128                 # 128x128 tiles, 8 K blocks (> stages), 2 sub-tiles per tile and
                                                        # 3 tiles per CTA (3 emulated SMs): staging buffers are re-used
                                                        # (the one-barrier epilogue is limited to the <= 128-wide tiles)
rng = np.random.RandomState(7)
x = from_bf16(to_bf16(rng.randn(B, H, W, cin))); w = from_bf16(to_bf16(rng.randn(cout, 1, 1, cin) / 22))
bias = (rng.randn(cout) * 0.1).astype(np.float32)
res = from_bf16(to_bf16(rng.randn(B, H, W, cout))); rb = to_bf16(res)
want = conv_ref(x, w, bias, 1, 1, res, True)
for late in (0, 1):
    lib.ssg_emu_set_async(late)
    y = np.zeros((B, H, W, cout), np.uint16); scratch = np.zeros(x.size + 64, np.uint16)
    xb, wb = to_bf16(x), to_bf16(w)
    rc = lib.ssg_op_conv(xb.ctypes.data, B, H, W, cin, 1, 1, wb.ctypes.data, bias.ctypes.data, cout, rb.ctypes.data, 1,
                         y.ctypes.data, scratch.ctypes.data, None)
    ok = rc == 0 and np.abs(from_bf16(y) - want).max() <= 0.01 * np.abs(want).max() + 0.02
    print("late" if late else "eager", "RIGHT" if ok else "WRONG")
"""


LATE_FAULT_CASES = [
    ("none", "0", {"eager": "RIGHT", "late": "RIGHT"}),
    ("none", "1", {"eager": "RIGHT", "late": "RIGHT"}),
    # the operand stage handed back to the TMA producer by a plain arrive instead of tcgen05.commit: the refill lands
    # before the MMAs that read the stage have executed
    ("stage_freed_early", "0", {"eager": "RIGHT", "late": "WRONG"}),
    # one-barrier epilogue without the wait on the previous TMA store: the staging buffer is rewritten under it
    ("epi2_no_store_wait", "1", {"eager": "RIGHT", "late": "WRONG"}),
    ("epi2_no_store_wait", "0", {"eager": "RIGHT", "late": "RIGHT"}),      # the fault sits in code EPI2 = 0 never runs
]


def test_late_completion_model_catches_injected_protocol_faults():
    """gemm_tc.cuh rebuilt with ONE deliberate protocol violation (build_emu.FAULTS): invisible when asynchronous work
    completes at issue, a wrong result when it completes as late as the protocol allows -- which is what makes "same
    bytes under both models" (tests above) evidence about the real kernels' synchronisation."""
    import build_emu
    from concurrent.futures import ThreadPoolExecutor
    for fault in sorted({c[0] for c in LATE_FAULT_CASES}):
        build_emu.build_tc(fault=None if fault == "none" else fault)          # before the subprocesses race for it
    paths = [ROOT, os.path.join(ROOT, "self-similarity-grouping_b200"), os.path.join(ROOT, "tests", "cpu_cuda"),
             os.path.join(ROOT, "tests")]
    with ThreadPoolExecutor(max_workers=min(5, os.cpu_count() or 1)) as pool:
        runs = [pool.submit(subprocess.run, [sys.executable, "-c", FAULT_SCRIPT % (paths,), fault], capture_output=True,
                            text=True, timeout=1200, env=dict(os.environ, SSG_CONV_EPI2=epi2))
                for fault, epi2, _ in LATE_FAULT_CASES]
        for (fault, epi2, expect), fut in zip(LATE_FAULT_CASES, runs):
            r = fut.result()
            assert r.returncode == 0, (fault, epi2, r.stdout + r.stderr)
            got = dict(l.split() for l in r.stdout.splitlines() if l.startswith(("eager", "late")))
            assert got == expect, (fault, epi2, r.stdout + r.stderr)
