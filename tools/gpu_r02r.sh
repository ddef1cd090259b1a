#!/bin/bash
# round 2, final single-GPU call: smoke, gpu test suite, the bench line (+ reference arm), ncu launch list of one full-size step
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02r_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_smoke.log; tail -n 2 gpurun_out/r02r_smoke.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02r_gpu_suite.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_gpu_suite.log; tail -n 3 gpurun_out/r02r_gpu_suite.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; tail -n 2 gpurun_out/r02r_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02r_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'embed',d['embed']['ms_per_step'], 'rerank',d['rerank']['ms_per_step'], 'frac',d['roofline']['frac'], 'e2e',d['e2e']['ms_per_step'],d['e2e']['value'], 'u8',d['e2e_u8']['ms_per_step'])
print(d['e2e_reference_api']); print(d['result']['labels_sha1'], d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02r_bench_reference.json 2> gpurun_out/r02r_bench_reference.err; cut -c1-400 gpurun_out/r02r_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02r_launches.csv \
    python bench.py --quick --steps 1 --warmup 0 > gpurun_out/r02r_launches.out 2>&1
ls -la gpurun_out | grep r02r
