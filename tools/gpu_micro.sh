#!/bin/bash
# last GPU seconds of the round: parity-plane stem staging (SSG_STEM_PLANES=1) -- bit identity against the plain variants, then A/B
set -u
mkdir -p gpurun_out
SSG_STEM_PLANES=1 timeout 40 python -m pytest tests/test_gpu_embed.py -x -q -k "variants" > gpurun_out/r01p_tests.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r01p_tests.log
SSG_STEM_PLANES=1 timeout 20 python bench.py --quick --steps 2 --warmup 1 --n 2048 2>/dev/null | tee gpurun_out/r01p_quick_planes.json | cut -c1-230
timeout 20 python bench.py --quick --steps 2 --warmup 1 --n 2048 2>/dev/null | tee gpurun_out/r01p_quick_default.json | cut -c1-230
