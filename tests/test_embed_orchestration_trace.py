"""CPU: the HOST orchestration of the embedding forward (csrc/embed.cu: which launcher runs on which buffers, in which
order) checked without a GPU through call traces (tests/cpu_cuda/embed_trace.py).

embed.cu contains no kernels; compiled against a logging stand-in for conv.cu and the CUDA runtime it yields one log
line per launch with every pointer as an arena offset.  Two facts are pinned:
  * the refactored forward (bottleneck block as run_block()) issues exactly the launches of the revision whose GPU parity
    tests last ran green on a B200 (1171b66) -- the default path is unchanged;
  * the L2-chunked schedule (SSG_L2_CHUNK, direct launches and CUDA-graph capture alike) is, by symbolic replay with one
    value id per image-pass and tensor, the same function of the same inputs as the default schedule: every pass goes
    through the same operators with the same weights, and no buffer is overwritten before its last reader."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "cpu_cuda"))
VALIDATED_REV = "1171b66"        # embed.cu as of the round's last full `pytest -m gpu` run on a B200 (r01n / r01p)

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


@pytest.fixture(scope="module")
def libs():
    import embed_trace as T
    head = T.build("head", T.head_source())
    try:
        old = T.build("validated", T.git_source(VALIDATED_REV))
    except subprocess.CalledProcessError:
        old = None                                   # no git history (e.g. an exported snapshot)
    return T, head, old


@pytest.mark.parametrize("n,num_split,flip", [(5, 2, 1), (8, 3, 1), (3, 1, 0)])
def test_default_forward_issues_the_validated_launch_sequence(libs, n, num_split, flip):
    T, head, old = libs
    if old is None:
        pytest.skip("revision %s not available" % VALIDATED_REV)
    new_t = T.run(head, n, num_split=num_split, flip=flip)
    old_t = T.run(old, n, num_split=num_split, flip=flip)
    assert len(new_t) > 50
    assert new_t == old_t


@pytest.mark.parametrize("chunk,graph", [(4, 0), (4, 1), (3, 0), (7, 1)])
def test_l2_chunked_schedule_is_the_same_function(libs, chunk, graph):
    T, head, _ = libs
    n = 5                                            # 10 image-passes: chunks of 4 / 3 / 7 leave ragged tails
    base = T.run(head, n)
    want = T.replay(base, n)
    assert not any(str(v).startswith("GARBAGE") for v in want) and len(set(want)) == 2 * n
    got_t = T.run(head, n, env={"SSG_L2_CHUNK": str(chunk), "SSG_L2_GRAPH": str(graph)})
    assert len(got_t) > len(base)                    # the chunk loop really ran
    assert T.replay(got_t, n) == want


def test_replay_notices_a_clobbered_buffer(libs):
    """The symbolic replay is not vacuous: redirect one layer-1 output onto its own input and the tail changes."""
    T, head, _ = libs
    base = T.run(head, 5)
    want = T.replay(base, 5)
    broken = list(base)
    idx = [i for i, l in enumerate(broken) if l.startswith("conv3x3")][0]
    x_off = [t for t in broken[idx].split() if t.startswith("x=")][0][2:]
    broken[idx] = " ".join(("y=" + x_off) if t.startswith("y=") else t for t in broken[idx].split())
    assert T.replay(broken, 5) != want
