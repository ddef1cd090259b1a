#!/bin/bash
# GPU-box profiling recipe (run under gpurun, ONE GPU).  usage: bash tools/profile.sh <tag> [launches] [full] [traffic]
#   launches : every launch of one full step with its device time (cold-cache, serialised: compare SHARES)
#   full     : ncu --set full of the distance GEMM, stem + layer 1, layer 4 and the re-rank side kernels
#   traffic  : DRAM bytes + time of every convolution GEMM launch of one embedding batch
#   traffic_warm : the same without cache flushes between launches, default schedule vs SSG_L2_CHUNK (SSG_CHUNK=32)
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
traffic_warm)
  # the same with the caches NOT flushed between launches (--cache-control none), once for the default schedule and once
  # for L2-resident chunks of layers 1-2 (direct launches: the chunk loop as a CUDA graph is one opaque launch to ncu):
  # what tools/l2_traffic_model.py predicts in its warm mode (profiles/r01_l2_chunk_model.md: 29.7 -> 9.7 GB per batch).
  # Three metrics fit one pass, so nothing is replayed and the cache state is the program's own.
  # launches per batch of 512 images (1024 image-passes): stem 1 + layers 1-2 21 per chunk + layers 3-4 27; the first
  # batch (target set) is skipped, the second (source set) is captured
  C=${SSG_CHUNK:-32}
  for v in "SSG_L2_CHUNK=0" "SSG_L2_CHUNK=$C SSG_L2_GRAPH=0"; do
    name=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
    case "$v" in "SSG_L2_CHUNK=0") S=49 ;; *) S=$((1 + (1024 + C - 1) / C * 21 + 27)) ;; esac
    env $v timeout ${LIM:-600} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --cache-control none -k regex:gemm_kernel -s $S -c $S --csv \
        --log-file gpurun_out/${TAG}_conv_traffic_warm_${name}.csv \
        python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_conv_traffic_warm_${name}.out 2>&1
  done ;;
esac
done
ls -la gpurun_out/ | grep "${TAG}_" | tail -12
