SSG_X=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm -s 46 -c 8 --csv --log-file gpurun_out/r02l_l1.csv python bench.py --quick --steps 1 --warmup 0 --n 512 > /dev/null 2>&1
grep gemm gpurun_out/r02l_l1.csv | awk -F'","' '{print $5, $NF}' | cut -c1-110
bash tools/gpu_multi.sh 4
