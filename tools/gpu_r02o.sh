#!/bin/bash
# round 2, call o (ONE GPU): stem without its zero kernel row, KHS pair kernel, 256-wide tiles for K = 128: bit-identity + A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_embed.py -q -x > gpurun_out/r02o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_tests.log; tail -n 3 gpurun_out/r02o_tests.log
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/r02o_ab_default.json 2> gpurun_out/r02o_ab_default.err
SSG_KHS_PAIR=1 timeout 200 python bench.py $Q > gpurun_out/r02o_ab_khs_pair.json 2> gpurun_out/r02o_ab_khs_pair.err
SSG_WIDE_K=128 timeout 200 python bench.py $Q > gpurun_out/r02o_ab_wide_k128.json 2> gpurun_out/r02o_ab_wide_k128.err
SSG_WIDE_K=128 SSG_KHS_PAIR=1 timeout 200 python bench.py $Q > gpurun_out/r02o_ab_both.json 2> gpurun_out/r02o_ab_both.err
for f in default khs_pair wide_k128 both; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/r02o_ab_%s.json'%sys.argv[1])); k=d['kernels_ms_per_step']
    print('%-10s step %.1f embed %.1f | 1x1 %.1f 3x3 %.1f stem %.1f'%(sys.argv[1],d['ms_per_step'],d['embed_ms'],k.get('conv1x1_tc',0),k.get('conv3x3_tc',0),k.get('conv_stem_tc',0)))
except Exception as e: print(sys.argv[1],'ERR',e)
PY
done; tail -n 2 gpurun_out/r02o_ab_*.err
