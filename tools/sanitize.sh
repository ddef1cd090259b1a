#!/bin/bash
# compute-sanitizer sweep on small shapes (run under gpurun): memcheck over every kernel family, racecheck on the
# shared-memory heavy ones.  Writes gpurun_out/sanitize_*.log; prints the error summaries.
mkdir -p gpurun_out
SEL='sqdist_exact or stages_against_oracle[64 or stages_against_oracle[257 or no_rerank or re_ranking_init or ties or argument_errors'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_rerank.py -x -q -k "$SEL" > gpurun_out/sanitize_rerank.log 2>&1
echo "memcheck rerank rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_rerank.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cluster.py -x -q -k "not full_size and not 2000 and not 1024" > gpurun_out/sanitize_cluster.log 2>&1
echo "memcheck cluster rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_cluster.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tensor.py -x -q -k "within_error_bound and (128-128 or 130-257) or equals_exact_mode and 257 or heavy_ties" > gpurun_out/sanitize_tensor.log 2>&1
echo "memcheck tensor rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_tensor.log | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_embed.py -x -q -k "conv_blocks or pooled_tail or fold_bn or (trunk_end_to_end and 2)" > gpurun_out/sanitize_embed.log 2>&1
echo "memcheck embed rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_embed.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_rerank.py tests/test_gpu_cluster.py -x -q -k "stages_against_oracle[64 or (dbscan_matches and 200) or (eps_matches and 300)" > gpurun_out/sanitize_race.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_race.log | tail -3
