#!/bin/bash
# round 2, call j (ONE GPU): programmatic dependent launch: bit-identity test + A/B; per-launch list of the final state
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_embed.py -q -x > gpurun_out/r02j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02j_tests.log; tail -n 3 gpurun_out/r02j_tests.log
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/r02j_ab_default.json 2> gpurun_out/r02j_ab_default.err
SSG_PDL=0 timeout 200 python bench.py $Q > gpurun_out/r02j_ab_no_pdl.json 2> gpurun_out/r02j_ab_no_pdl.err
cat gpurun_out/r02j_ab_*.json; tail -n 2 gpurun_out/r02j_ab_*.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
    -k regex:gemm -s 46 -c 46 --csv --log-file gpurun_out/r02j_conv_traffic.csv \
    python bench.py --quick --steps 1 --warmup 0 --n 512 > gpurun_out/r02j_conv_traffic.out 2>&1
ls -la gpurun_out | grep r02j
