#!/bin/bash
# round 2, call d: ncu --set full of the stem + layer-1 + first layer-2 launches (final state: chained kernel, KHS with
# resident weights), and the per-launch time / DRAM-byte list of one whole embedding batch
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 46 -c 12 -o gpurun_out/r02d_l1l2 \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/r02d_l1l2.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
    -k regex:gemm_kernel -s 46 -c 46 --csv --log-file gpurun_out/r02d_conv_traffic.csv \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/r02d_conv_traffic.out 2>&1
ls -la gpurun_out | grep r02d
