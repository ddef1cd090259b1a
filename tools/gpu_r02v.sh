#!/bin/bash
# round 2, closing single-GPU call after the late additions (metrics kernel, training-side convolutions): smoke, the whole
# gpu test suite, the bench line (+ fine-tune step) and the reference arm
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02v_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_smoke.log; tail -n 2 gpurun_out/r02v_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02v_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_gpu_suite.log; tail -n 4 gpurun_out/r02v_gpu_suite.log
timeout 600 python bench.py --steps 5 --warmup 3 --finetune-step > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; tail -n 2 gpurun_out/r02v_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02v_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'embed',d['embed']['ms_per_step'], 'rerank',d['rerank']['ms_per_step'], 'frac',d['roofline']['frac'], 'e2e',d['e2e']['ms_per_step'],d['e2e']['value'], 'u8',d['e2e_u8']['ms_per_step'])
print(d['e2e_reference_api']['ms_per_step'], d['e2e_reference_api']['labels_equal_device_resident_path']); print(d['result']['labels_sha1'], d['clocks'], d['parity_gate']['ok'])
print(d['finetune_step'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02v_bench_reference.json 2> gpurun_out/r02v_bench_reference.err; cut -c1-300 gpurun_out/r02v_bench_reference.json
