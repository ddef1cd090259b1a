#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent.
#   gpurun --timeout 900 -- 'bash tools/gpu_next.sh'            (one GPU, ~6 min)
# 1. plain-C harnesses on the C ABI (seconds each, no Python start-up)   2. pytest -m gpu_next
# 3. A/B timings of the opt-in variants against the defaults (bench.py --quick: CUDA-event timers, no CPU legs)
# Each step has its own time limit; results land in gpurun_out/next_*.
mkdir -p gpurun_out
make -C tests/c > gpurun_out/next_make.log 2>&1
timeout 60 tests/c/_build/shard_check 4000 8 > gpurun_out/next_shard_check.log 2>&1; echo "rc=$?" >> gpurun_out/next_shard_check.log
timeout 60 tests/c/_build/sparse_check 4000 128 > gpurun_out/next_sparse_check.log 2>&1; echo "rc=$?" >> gpurun_out/next_sparse_check.log
# embedding variants, byte-compared against the default through the plain-C harness (seconds each)
D=tests/c/_build/embed_dump
timeout 120 $D /tmp/emb_default.bin > gpurun_out/next_embed_dump.log 2>&1
for v in "SSG_CONV_EPI2=1" "SSG_L2_CHUNK=8" "SSG_L2_CHUNK=10 SSG_L2_GRAPH=0" "SSG_L2_CHUNK=32" "SSG_CONV_EPI2=1 SSG_L2_CHUNK=16"; do
  env $v timeout 120 $D /tmp/emb_variant.bin >> gpurun_out/next_embed_dump.log 2>&1
  if cmp -s /tmp/emb_default.bin /tmp/emb_variant.bin; then echo "$v: identical to the default" >> gpurun_out/next_embed_dump.log
  else echo "$v: DIFFERS from the default" >> gpurun_out/next_embed_dump.log; fi
done
timeout 600 python -m pytest tests -m gpu_next -q -x > gpurun_out/next_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/next_pytest.log
Q="--quick --steps 2 --warmup 1"
timeout 200 python bench.py $Q > gpurun_out/next_ab_default.json 2> gpurun_out/next_ab_default.err
timeout 200 python bench.py $Q --sparse-finish > gpurun_out/next_ab_sparse.json 2> gpurun_out/next_ab_sparse.err
SSG_DIST_SYM=1 timeout 200 python bench.py $Q > gpurun_out/next_ab_distsym.json 2> gpurun_out/next_ab_distsym.err
SSG_DIST_SYM=1 timeout 200 python bench.py $Q --sparse-finish > gpurun_out/next_ab_distsym_sparse.json 2> gpurun_out/next_ab_distsym_sparse.err
SSG_PAIR_VEC8=1 timeout 200 python bench.py $Q > gpurun_out/next_ab_pairvec8.json 2> gpurun_out/next_ab_pairvec8.err
SSG_CONV_EPI2=1 timeout 200 python bench.py $Q > gpurun_out/next_ab_epi2.json 2> gpurun_out/next_ab_epi2.err
for c in 8 16 32 48; do
  SSG_L2_CHUNK=$c timeout 200 python bench.py $Q > gpurun_out/next_ab_l2chunk$c.json 2> gpurun_out/next_ab_l2chunk$c.err
done
# chunks and the one-barrier epilogue together: once the HBM bytes are gone the epilogue is the next limit
# (profiles/r01_l2_chunk_model.md section 3)
SSG_L2_CHUNK=32 SSG_CONV_EPI2=1 timeout 200 python bench.py $Q > gpurun_out/next_ab_l2chunk32_epi2.json 2> gpurun_out/next_ab_l2chunk32_epi2.err
tail -n 3 gpurun_out/next_*.log; cat gpurun_out/next_ab_*.json
# multi-GPU (separate call, gpurun --gpus 2/8):
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus 8 --steps 3 --warmup 3 [--sparse-finish | --shard-finish]
# L2-chunk traffic against the LRU model (profiles/r01_l2_chunk_model.md), ~2 min more:
#   bash tools/profile.sh r02a traffic_warm        (SSG_CHUNK=32 by default; DRAM bytes per launch, caches not flushed)
