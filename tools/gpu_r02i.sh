#!/bin/bash
# round 2, call i (ONE GPU): why the reference-API leg's eps differs; quick A/B after the pooled-tail change
mkdir -p gpurun_out
timeout 600 python tools/debug_api_vs_device.py 16702 > gpurun_out/r02i_debug.json 2> gpurun_out/r02i_debug.err; tail -n 3 gpurun_out/r02i_debug.err; cat gpurun_out/r02i_debug.json
timeout 200 python bench.py --quick --steps 2 --warmup 1 > gpurun_out/r02i_ab_default.json 2> gpurun_out/r02i_ab_default.err; cat gpurun_out/r02i_ab_default.json
timeout 300 python -m pytest tests/test_gpu_embed.py -q -x 2>&1 | tail -n 3
