#!/bin/bash
# round 2, call c: chained layer-1 kernel + resident-weight KHS: bit-identity test, then A/B timings (bench --quick)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_embed.py tests/test_gpu_whole_path.py -q -x > gpurun_out/r02c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_tests.log
tail -n 5 gpurun_out/r02c_tests.log
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/r02c_ab_default.json 2> gpurun_out/r02c_ab_default.err
SSG_CONV_CHAIN=0 timeout 200 python bench.py $Q > gpurun_out/r02c_ab_nochain.json 2> gpurun_out/r02c_ab_nochain.err
SSG_KHS_BRES=0 timeout 200 python bench.py $Q > gpurun_out/r02c_ab_khs_streamed.json 2> gpurun_out/r02c_ab_khs_streamed.err
SSG_CONV_CHAIN=0 SSG_KHS_BRES=0 timeout 200 python bench.py $Q > gpurun_out/r02c_ab_round1.json 2> gpurun_out/r02c_ab_round1.err
SSG_CONV_EPI2=1 timeout 200 python bench.py $Q > gpurun_out/r02c_ab_epi2.json 2> gpurun_out/r02c_ab_epi2.err
cat gpurun_out/r02c_ab_*.json; tail -n 2 gpurun_out/r02c_ab_*.err
