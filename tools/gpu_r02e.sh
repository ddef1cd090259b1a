#!/bin/bash
# round 2, call e: deeper pipelines (VAR_NORES, KHS 5 stages, chained kernel's residual ring) -- tests, A/B, then the
# full single-GPU bench line with its new legs (parity gate, uint8 e2e, reference-API e2e) and the reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_embed.py tests/test_gpu_whole_path.py -q -x > gpurun_out/r02e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_tests.log
tail -n 4 gpurun_out/r02e_tests.log
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/r02e_ab_default.json 2> gpurun_out/r02e_ab_default.err
SSG_CONV_NORES=0 timeout 200 python bench.py $Q > gpurun_out/r02e_ab_no_nores.json 2> gpurun_out/r02e_ab_no_nores.err
cat gpurun_out/r02e_ab_*.json; tail -n 2 gpurun_out/r02e_ab_*.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -n 3 gpurun_out/r02e_bench.err; cat gpurun_out/r02e_bench.json | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_reference.json 2> gpurun_out/r02e_bench_reference.err
cat gpurun_out/r02e_bench_reference.json | cut -c1-1500
