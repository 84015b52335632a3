#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_r1d.log
( FOKL_GRAM_WARPS=16 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_r1d_w16.log
( FOKL_GRAM_WARPS=8 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_r1d_w8.log
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 11 -c 2 -f -o gpurun_out/prof_gram4 $B0 > gpurun_out/prof_gram4.log 2>&1
cat gpurun_out/pytest_r1d.log; cat gpurun_out/bench_r1d_w16.log; cat gpurun_out/bench_r1d_w8.log
