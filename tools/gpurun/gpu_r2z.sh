#!/bin/bash
mkdir -p gpurun_out
for mp in 256 160; do
( FOKL_B200_NESTED_MIN_P=$mp timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2z_bench_cfg5_$mp.log
echo "cfg5 nested_min_p=$mp"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z_bench_cfg5_$mp.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2z_bench_cfg5_$mp.log
done
for mp in 190 120; do
( FOKL_B200_NESTED_MIN_P=$mp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2z_bench_cfg4_$mp.log
echo "cfg4 nested_min_p=$mp"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2z_bench_cfg4_$mp.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2z_bench_cfg4_$mp.log
done
