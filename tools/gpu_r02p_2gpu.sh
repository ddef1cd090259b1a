# de-risk configs[4] on 2 GPUs at N = 60 000 (pattern tensor > 2^31 bytes, 14 GB of final_dist rows per rank)
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --num-images 60000 --shard-finish --steps 1 --warmup 3 --finetune-step > gpurun_out/r02p_n60000_2gpu.json 2> gpurun_out/r02p_n60000_2gpu.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|vectorized_gather" gpurun_out/r02p_n60000_2gpu.err | tail -n 12 | cut -c1-300
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02p_n60000_2gpu.json') if l.startswith('{')][-1]; r=d['result']
    print('ms %.1f e2e %.1f embed %s rerank %.1f clusters %s eps %s'%(d['ms_per_step'],d['e2e']['ms_per_step'],(d['embed'] or {}).get('ms_per_step'),d['rerank']['ms_per_step'],r['clusters'],r['eps']))
    print(d.get('finetune_step'))
except Exception as e: print('ERR',e)
PY
