#!/bin/bash
# round 2, call h (ONE GPU): staged pair_exact (v4): re-rank / cluster / variant tests + A/B; full bench line; single-GPU
# lines of configs[2] and configs[3] (label checksums for the multi-GPU runs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_tensor.py tests/test_gpu_cluster.py tests/test_gpu_next_variants.py tests/test_gpu_api_rows.py -q -x > gpurun_out/r02h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_tests.log
tail -n 4 gpurun_out/r02h_tests.log
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/r02h_ab_default.json 2> gpurun_out/r02h_ab_default.err
SSG_PAIR_STAGED=0 timeout 200 python bench.py $Q > gpurun_out/r02h_ab_pair_exact_v3.json 2> gpurun_out/r02h_ab_pair_exact_v3.err
cat gpurun_out/r02h_ab_*.json
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; tail -n 2 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['embed']['ms_per_step'], d['rerank']['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e_reference_api'], d['result']['labels_sha1'])
PY
timeout 600 python bench.py --banks 4 --steps 3 --warmup 3 --no-cpu-baseline --no-u8 --no-reference-api > gpurun_out/r02h_config2_1gpu.json 2> gpurun_out/r02h_config2_1gpu.err; tail -n 2 gpurun_out/r02h_config2_1gpu.err
for rho in 0.0008 0.0016 0.0032; do
  timeout 600 python bench.py --n 36411 --features-only --rho $rho --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_config3_1gpu_rho$rho.json 2> gpurun_out/r02h_config3_1gpu_rho$rho.err; tail -n 2 gpurun_out/r02h_config3_1gpu_rho$rho.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02h_config*.json')):
    try:
        d=json.load(open(f)); print(f, d['ms_per_step'], d['rerank']['ms_per_step'], d['result']['clusters'], d['result']['labels_sha1'])
    except Exception as e: print(f,'ERR',e)
PY
