#!/bin/bash
# round 2: the fine-tune step captured in a CUDA graph (test + timing)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_ops.py -q -s -k "graphed" > gpurun_out/r02x_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02x_tests.log; grep -v "^$" gpurun_out/r02x_tests.log | tail -n 25
timeout 600 python bench.py --n 2048 --steps 1 --warmup 3 --finetune-step --no-cpu-baseline --no-reference-api --no-u8 \
    > gpurun_out/r02x_bench_finetune.json 2> gpurun_out/r02x_bench_finetune.err
tail -n 3 gpurun_out/r02x_bench_finetune.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02x_bench_finetune.json'))
    print(json.dumps(d.get('finetune_step'), indent=1))
except Exception as e:
    print('no bench line', e)
PY
