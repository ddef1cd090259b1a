#!/bin/bash
# GPU-box profiling recipe (run under gpurun).  $1 = tag (e.g. r01d), $2 = "launches" to also take the launch list
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
if [ "${2:-}" = "launches" ]; then
  # every launch of one full step with its device time (cold-cache, serialised: compare SHARES); ~25 min
  timeout 1700 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
      python bench.py --quick --steps 1 --warmup 0 > gpurun_out/${TAG}_launches.out 2>&1
fi
# full-set capture: the distance GEMM (2 launches) and one whole batch of the convolution GEMMs (53 launches)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 2 -o gpurun_out/${TAG}_gemm_dist \
    python bench.py --quick --steps 1 --warmup 0 --features-only > gpurun_out/${TAG}_gemm_dist.out 2>&1
# (reports are ~1.4 MB per launch and gpurun_out/ is capped at 64 MiB: stem + layer1 + start of layer2, then layer4)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 53 -c 16 -o gpurun_out/${TAG}_gemm_conv_l1 \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_gemm_conv_l1.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 96 -c 10 -o gpurun_out/${TAG}_gemm_conv_l4 \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/${TAG}_gemm_conv_l4.out 2>&1
ls -la gpurun_out/ | tail -8
# re-rank side kernels (one bank): select / rescoring / eps / Jaccard / DBSCAN scans
if [ "${3:-}" = "rerank" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'row_select|pair_exact|eps_hist|eps_gather|jaccard_final|db_count|db_fill' -c 14 -o gpurun_out/${TAG}_rerank \
    python bench.py --quick --steps 1 --warmup 0 --features-only --banks 1 > gpurun_out/${TAG}_rerank.out 2>&1
fi
