#!/bin/bash
# One-call GPU validation (run under gpurun, ONE GPU): the whole -m gpu suite, the default bench line, smoke(), the
# ncu launch list of one full step and the conv DRAM-traffic capture.  Every step has its own timeout and logs under
# gpurun_out/<tag>_*.
set -u
TAG=${1:-r01n}
mkdir -p gpurun_out
T0=$(date +%s)
step() { echo "== $1  (t=$(( $(date +%s)-T0 ))s)"; }
step "gpu suite"; timeout 420 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "rc=$?"; tail -14 gpurun_out/${TAG}_gpu_tests.log
step "bench"; timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"; head -c 700 gpurun_out/${TAG}_bench.json; echo
step "smoke"; timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
step "launches"; SSG_PROFILE_TIMEOUT=600 bash tools/profile.sh ${TAG} launches > /dev/null 2>&1; echo "rc=$?"
step "traffic"; SSG_PROFILE_TIMEOUT=150 bash tools/profile.sh ${TAG} traffic > /dev/null 2>&1; echo "rc=$?"
for b in 256 1024; do
  step "quick batch $b"; timeout 100 python bench.py --quick --steps 2 --warmup 1 --n 4096 --batch $b 2>/dev/null | cut -c1-330
done
step "done"
