#!/bin/bash
# round 2, validation of the late additions: retrieval-metrics kernel, training-side convolution operators, fine-tune step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_cluster.py tests/test_gpu_api_rows.py -q -s -x > gpurun_out/r02s_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02s_tests.log; tail -n 25 gpurun_out/r02s_tests.log
timeout 600 python bench.py --n 2048 --steps 1 --warmup 3 --finetune-step --no-cpu-baseline --no-reference-api --no-u8 \
    > gpurun_out/r02s_bench_finetune.json 2> gpurun_out/r02s_bench_finetune.err
tail -n 3 gpurun_out/r02s_bench_finetune.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02s_bench_finetune.json'))
    print(json.dumps(d.get('finetune_step'), indent=1)); print('ms_per_step', d['ms_per_step'], d['parity_gate'])
except Exception as e:
    print('no bench line', e)
PY
