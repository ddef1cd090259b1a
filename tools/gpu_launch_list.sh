set -u
mkdir -p gpurun_out
timeout 60 python bench.py --quick --steps 1 --warmup 1 --n 2048 > gpurun_out/r01o_quick_n2048.json 2>/dev/null; echo "rc=$?"; cut -c1-400 gpurun_out/r01o_quick_n2048.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01o_launches.csv python bench.py --quick --steps 1 --warmup 0 --n 2048 > gpurun_out/r01o_launches.out 2>&1; echo "rc=$?"; wc -l gpurun_out/r01o_launches.csv
