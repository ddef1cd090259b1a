#!/bin/bash
# round 2, call b: the new API-row tests, the whole-path parity gate, every gpu_next test (no -x), then the gpu suite
mkdir -p gpurun_out
make -C tests/c > gpurun_out/r02b_make.log 2>&1
timeout 600 python -m pytest tests/test_gpu_api_rows.py tests/test_gpu_whole_path.py -q -s > gpurun_out/r02b_new_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_new_tests.log
timeout 900 python -m pytest tests -m gpu_next -q > gpurun_out/r02b_gpu_next.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_gpu_next.log
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_api_rows.py --deselect tests/test_gpu_whole_path.py > gpurun_out/r02b_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_gpu_suite.log
tail -n 25 gpurun_out/r02b_new_tests.log; tail -n 40 gpurun_out/r02b_gpu_next.log; tail -n 8 gpurun_out/r02b_gpu_suite.log
