#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_r1c.log
( timeout 600 python tools/explore.py --cfg cfg4 --repeat 3 --resident 1 --cprofile 1 2>&1 | tail -90 ) > gpurun_out/explore_resident.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/bench_r1c.log
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 11 -c 2 -f -o gpurun_out/prof_gram3 $B0 > gpurun_out/prof_gram3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:basis_kernel -s 3 -c 4 -f -o gpurun_out/prof_basis3 $B0 > gpurun_out/prof_basis3.log 2>&1
cat gpurun_out/pytest_r1c.log; cat gpurun_out/bench_r1c.log; tail -50 gpurun_out/explore_resident.log
