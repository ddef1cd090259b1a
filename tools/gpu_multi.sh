#!/bin/bash
# Multi-GPU runs of BASELINE.json's configs (gpurun --gpus N -- 'bash tools/gpu_multi.sh N').  Every run leaves one JSON line
# under gpurun_out/r02_multi_*; labels are compared across GPU counts through result.labels_sha1 (tools/gpu_r02g.sh holds
# the single-GPU lines).
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { # name, timeout, args...
  name=$1; lim=$2; shift 2
  timeout $lim $TR bench.py --gpus $N "$@" > gpurun_out/r02_multi_${name}_${N}gpu.json 2> gpurun_out/r02_multi_${name}_${N}gpu.err
  echo "== $name rc=$?"; tail -n 2 gpurun_out/r02_multi_${name}_${N}gpu.err | cut -c1-300
  python - "$name" "$N" <<'PY'
import json,sys
f='gpurun_out/r02_multi_%s_%sgpu.json'%(sys.argv[1],sys.argv[2])
try:
    d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]; r=d['result']
    print(sys.argv[1], 'value %.1f ms %.1f e2e_ms %.1f embed %s rerank %.1f clusters %s sha1 %s'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],(d['embed'] or {}).get('ms_per_step'),d['rerank']['ms_per_step'],r['clusters'],r['labels_sha1']))
    if d.get('finetune_step'): print('finetune', d['finetune_step'])
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
case "$N" in
2)
  run config1 600 --steps 3 --warmup 3
  run config2 600 --banks 4 --steps 3 --warmup 3 ;;
4)
  run config1 600 --steps 3 --warmup 3
  for rho in 0.0008 0.0016 0.0032; do run config3_rho$rho 600 --num-images 36411 --features-only --rho $rho --steps 2 --warmup 3; done ;;
8)
  run config1 600 --steps 3 --warmup 3
  run config1_shard_finish 600 --steps 3 --warmup 3 --shard-finish
  run config4 1200 --num-images 126441 --shard-finish --steps 1 --warmup 3 --finetune-step ;;
esac
