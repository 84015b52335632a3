#!/bin/bash
mkdir -p gpurun_out
for s in 1 0; do
echo "FOKL_EIGB_SORT=$s"
( FOKL_EIGB_SORT=$s timeout 300 python tools/eig_big_check.py 1024,2072 1 gauss,spline 2>&1 | grep -v "Warn\|chain" )
done > gpurun_out/r2t_sort.log; cat gpurun_out/r2t_sort.log
( FOKL_EIGB_MIN_P=2 timeout 120 python tools/eig_big_check.py 17,64,130,300 1 gauss 2>&1 | grep -v "Warn\|chain" ) > gpurun_out/r2t_small.log; cat gpurun_out/r2t_small.log
for s in 1 0; do
( FOKL_EIGB_SORT=$s timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2t_bench_cfg5_$s.log
echo "cfg5 sort=$s"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2t_bench_cfg5_$s.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2t_bench_cfg5_$s.log
done
