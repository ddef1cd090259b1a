# de-risk the 8-GPU call on 2 GPUs: row-sharded finish, sparse owners, a config-4-like run with the fine-tune step
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { name=$1; shift
  timeout 600 $TR bench.py --gpus $N "$@" > gpurun_out/r02n_${name}_${N}gpu.json 2> gpurun_out/r02n_${name}_${N}gpu.err
  echo "== $name rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/r02n_${name}_${N}gpu.err | tail -n 4 | cut -c1-300
  python - "$name" <<'PY'
import json,sys
f='gpurun_out/r02n_%s_2gpu.json'%sys.argv[1]
try:
    d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]; r=d['result']
    print(sys.argv[1], 'ms %.1f e2e %.1f embed %s rerank %.1f clusters %s sha1 %s'%(d['ms_per_step'],d['e2e']['ms_per_step'],(d['embed'] or {}).get('ms_per_step'),d['rerank']['ms_per_step'],r['clusters'],r['labels_sha1']))
    if d.get('finetune_step'): print('finetune', d['finetune_step'])
except Exception as e: print(sys.argv[1],'ERR',e)
PY
}
run shard_finish --steps 2 --warmup 3 --shard-finish
run sparse_owner --steps 2 --warmup 3 --sparse-finish
run like_config4 --num-images 30000 --shard-finish --steps 1 --warmup 3 --finetune-step
