#!/bin/bash
# GPU-box profiling recipe (run under gpurun).  $1 = tag (e.g. r01b)
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
# 1) launch list of one full step (cold-cache, serialised: compare SHARES)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --quick --steps 1 --warmup 0 > gpurun_out/${TAG}_launches.out 2>&1
# 2) full-set capture of the two GEMM flavours (3 launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:EpiDist -c 2 -o gpurun_out/${TAG}_gemm_dist \
    python bench.py --quick --steps 1 --warmup 0 --features-only > gpurun_out/${TAG}_gemm_dist.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:StagedEpi -s 40 -c 6 -o gpurun_out/${TAG}_gemm_conv \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_gemm_conv.out 2>&1
ls -la gpurun_out/ | tail -12
