#!/bin/bash
# One-call GPU validation (run under gpurun, ONE GPU): new-row tests, the default bench line, the whole -m gpu suite,
# the conv DRAM-traffic capture and smoke().  Every step has its own timeout and logs under gpurun_out/<tag>_*.
set -u
TAG=${1:-r01j}
mkdir -p gpurun_out
T0=$(date +%s)
step() { echo "== $1  (t=$(( $(date +%s)-T0 ))s)"; }
step "new tests"; timeout 240 python -m pytest tests/test_gpu_triplet.py "tests/test_gpu_embed.py::test_uint8_pixels_give_bit_identical_features" -x -q > gpurun_out/${TAG}_newtests.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/${TAG}_newtests.log
step "bench"; timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; head -c 1200 gpurun_out/${TAG}_bench.json; echo
step "gpu suite"; timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "rc=$?"; tail -22 gpurun_out/${TAG}_gpu_tests.log
step "traffic"; timeout 150 bash tools/profile.sh ${TAG} traffic > /dev/null 2>&1; echo "rc=$?"
step "smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
step "done"
