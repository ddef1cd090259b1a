#!/bin/bash
# round 2, call f: (1) the two-CTA (cta_group::2) kernel: byte comparison of the embedding against the default through the
# plain-C harness, each run under its own short timeout; (2) tensor vs exact distance mode on real trunk features
mkdir -p gpurun_out
make -C tests/c > gpurun_out/r02f_make.log 2>&1
D=tests/c/_build/embed_dump
timeout 120 $D /tmp/emb_default.bin > gpurun_out/r02f_pair.log 2>&1
for v in "SSG_CONV_PAIR=1" "SSG_CONV_PAIR=2" "SSG_CONV_PAIR=3"; do
  env $v timeout 60 $D /tmp/emb_variant.bin >> gpurun_out/r02f_pair.log 2>&1; echo "$v rc=$?" >> gpurun_out/r02f_pair.log
  if cmp -s /tmp/emb_default.bin /tmp/emb_variant.bin; then echo "$v: identical to the default" >> gpurun_out/r02f_pair.log
  else echo "$v: DIFFERS from the default" >> gpurun_out/r02f_pair.log; fi
  rm -f /tmp/emb_variant.bin
done
cat gpurun_out/r02f_pair.log
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 300 python tools/check_modes_on_cnn_features.py 4096 > gpurun_out/r02f_modes.json 2> gpurun_out/r02f_modes.err; tail -n 3 gpurun_out/r02f_modes.err; cat gpurun_out/r02f_modes.json
Q="--quick --steps 2 --warmup 1"
for v in 1 2 3; do
  SSG_CONV_PAIR=$v timeout 120 python bench.py $Q > gpurun_out/r02f_ab_pair$v.json 2> gpurun_out/r02f_ab_pair$v.err; tail -n 1 gpurun_out/r02f_ab_pair$v.err; cat gpurun_out/r02f_ab_pair$v.json
done
