#!/bin/bash
# session 4, call h: kill loop -- ncu of the single-CTA kernel on a 197-column model, and the whole-device kernel from 150 columns
mkdir -p gpurun_out
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kill_loop_kernel -s 6 -c 1 -f -o gpurun_out/r4h_prof_kill $B0 > gpurun_out/r4h_prof_kill.log 2>&1
tail -2 gpurun_out/r4h_prof_kill.log | cut -c1-200
for mp in 150 100; do
( FOKL_KILL_BIG_MIN_P=$mp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4h_bench_cfg4_big$mp.log
echo "kill big from $mp"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4h_bench_cfg4_big$mp.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4h_bench_cfg4_big$mp.log
done
