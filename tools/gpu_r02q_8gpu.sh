# configs[4] on 8 GPUs: N = 126 441 (MSMT17 shape), full cycle sharded with the row-sharded finish + one fine-tune step
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 1200 $TR bench.py --gpus $N --num-images 126441 --shard-finish --steps 1 --warmup 3 --finetune-step > gpurun_out/r02_multi_config4_8gpu.json 2> gpurun_out/r02_multi_config4_8gpu.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|vectorized_gather" gpurun_out/r02_multi_config4_8gpu.err | tail -n 12 | cut -c1-300
python - <<'PY'
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02_multi_config4_8gpu.json') if l.startswith('{')][-1]; r=d['result']
    print('value %.1f ms %.1f e2e %.1f embed %s rerank %.1f clusters %s eps %s kept %s'%(d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],(d['embed'] or {}).get('ms_per_step'),d['rerank']['ms_per_step'],r['clusters'],r['eps'],r['kept_images']))
    print(d.get('finetune_step'))
except Exception as e: print('ERR',e)
PY
nvidia-smi --query-gpu=memory.used --format=csv | head -3
