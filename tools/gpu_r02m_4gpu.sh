N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for rho in 0.0008 0.0016 0.0032; do
  timeout 600 $TR bench.py --gpus $N --num-images 36411 --features-only --rho $rho --steps 2 --warmup 3 > gpurun_out/r02_multi_config3_rho${rho}_${N}gpu.json 2> gpurun_out/r02_multi_config3_rho${rho}_${N}gpu.err
  echo "rho $rho rc=$?"; tail -n 2 gpurun_out/r02_multi_config3_rho${rho}_${N}gpu.err | cut -c1-300
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_multi_config3_*_4gpu.json')):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith('{')][-1]; r=d['result']
        print(f, 'ms %.1f e2e %.1f rerank %.1f clusters %s sha1 %s'%(d['ms_per_step'],d['e2e']['ms_per_step'],d['rerank']['ms_per_step'],r['clusters'],r['labels_sha1']))
    except Exception as e: print(f,'ERR',e)
PY
