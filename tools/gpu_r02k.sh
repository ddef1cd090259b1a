#!/bin/bash
# round 2, call k (ONE GPU): chained kernel: residual ring of 5 vs 3 slots (live A/B + per-launch times under ncu)
mkdir -p gpurun_out
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/r02k_ab_ring5.json 2> gpurun_out/r02k_ab_ring5.err
SSG_CHAIN_RING=3 timeout 200 python bench.py $Q > gpurun_out/r02k_ab_ring3.json 2> gpurun_out/r02k_ab_ring3.err
SSG_CHAIN_RING=4 timeout 200 python bench.py $Q > gpurun_out/r02k_ab_ring4.json 2> gpurun_out/r02k_ab_ring4.err
cat gpurun_out/r02k_ab_*.json
for r in 0 3; do
SSG_CHAIN_RING=$r timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm -s 46 -c 8 --csv --log-file gpurun_out/r02k_l1_ring$r.csv \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > /dev/null 2>&1
grep gemm gpurun_out/r02k_l1_ring$r.csv | awk -F'","' '{print $5, $NF}' | cut -c1-120
done
