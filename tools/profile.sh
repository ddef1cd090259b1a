#!/bin/bash
# GPU-box profiling recipe (run under gpurun, ONE GPU).  usage: bash tools/profile.sh <tag> [launches] [full] [traffic]
#   launches : every launch of one full step with its device time (cold-cache, serialised: compare SHARES)
#   full     : ncu --set full of the distance GEMM, stem + layer 1, layer 4 and the re-rank side kernels
#   traffic  : DRAM bytes + time of every convolution GEMM launch of one embedding batch
# Reports are 1.4-3 MB per launch and gpurun_out/ is capped at 64 MiB IN TOTAL per call (everything is dropped beyond
# that), hence the small launch counts: 2 + 8 + 5 + 8 launches ~ 45 MB.
set -u
# per-step limit: override with SSG_PROFILE_TIMEOUT (seconds).  Do NOT wrap this script in an outer `timeout`: killing
# the shell leaves ncu and its child running on the GPU (that contaminated the r01n batch sweep).
LIM=${SSG_PROFILE_TIMEOUT:-}
TAG=${1:-r01}
shift
mkdir -p gpurun_out
for what in "$@"; do
case "$what" in
launches)
  timeout ${LIM:-1500} ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --quick --steps 1 --warmup 0 > gpurun_out/${TAG}_launches.out 2>&1 ;;
full)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o gpurun_out/${TAG}_gemm_dist \
      python bench.py --quick --steps 1 --warmup 0 --features-only > gpurun_out/${TAG}_gemm_dist.out 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 49 -c 8 -o gpurun_out/${TAG}_gemm_conv_l1 \
      python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_gemm_conv_l1.out 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 89 -c 5 -o gpurun_out/${TAG}_gemm_conv_l4 \
      python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_gemm_conv_l4.out 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:'row_select|pair_exact|eps_hist|eps_gather|jaccard_final|db_count|db_fill' -c 8 -o gpurun_out/${TAG}_rerank \
      python bench.py --quick --steps 1 --warmup 0 --features-only --banks 1 > gpurun_out/${TAG}_rerank.out 2>&1 ;;
traffic)
  timeout ${LIM:-600} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:gemm_kernel -s 49 -c 49 --csv --log-file gpurun_out/${TAG}_conv_traffic.csv \
      python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_conv_traffic.out 2>&1 ;;
esac
done
ls -la gpurun_out/ | grep "${TAG}_" | tail -12
