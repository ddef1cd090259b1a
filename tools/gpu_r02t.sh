#!/bin/bash
# round 2: second validation of the late additions (network-level fine-tune test on a smooth loss, metrics kernel, API rows)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_cluster.py tests/test_gpu_api_rows.py -q -s > gpurun_out/r02t_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r02t_tests.log; tail -n 30 gpurun_out/r02t_tests.log
