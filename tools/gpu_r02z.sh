#!/bin/bash
# round 2, last call: smoke + the whole gpu suite on the final tree
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/r02z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_smoke.log; tail -n 2 gpurun_out/r02z_smoke.log
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r02z_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_gpu_suite.log; tail -n 4 gpurun_out/r02z_gpu_suite.log
