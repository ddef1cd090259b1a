#!/bin/bash
# A/B of the conv variants on one GPU (run under gpurun): tests with the defaults, then quick benches with and
# without the env switches, then one ncu --set full capture of the stem kernel.
set -u
TAG=${1:-r01k}
mkdir -p gpurun_out
T0=$(date +%s)
step() { echo "== $1  (t=$(( $(date +%s)-T0 ))s)"; }
step "embed+cluster tests"; timeout 300 python -m pytest tests/test_gpu_embed.py tests/test_gpu_cluster.py -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/${TAG}_tests.log
step "quick default"; timeout 120 python bench.py --quick --steps 2 --warmup 1 --n 4096 > gpurun_out/${TAG}_quick_default.json 2> gpurun_out/${TAG}_quick_default.err; echo "rc=$?"; cat gpurun_out/${TAG}_quick_default.json
step "quick old variants"; SSG_STEM_BRES=0 SSG_STEM_POOL=0 SSG_CONV_BN256_RES=0 timeout 120 python bench.py --quick --steps 2 --warmup 1 --n 4096 > gpurun_out/${TAG}_quick_old.json 2> gpurun_out/${TAG}_quick_old.err; echo "rc=$?"; cat gpurun_out/${TAG}_quick_old.json
step "ncu stem"; timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 1 -o gpurun_out/${TAG}_stem python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_stem.out 2>&1; echo "rc=$?"
step "traffic"; SSG_PROFILE_TIMEOUT=150 bash tools/profile.sh ${TAG} traffic > /dev/null 2>&1; echo "rc=$?"
step "done"
