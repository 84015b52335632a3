#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gram2.log
( timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -2 ) > gpurun_out/bench_gram2.log
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 8 -c 6 -f -o gpurun_out/prof_gram2 $B0 > gpurun_out/prof_gram2.log 2>&1
cat gpurun_out/pytest_gram2.log; cat gpurun_out/bench_gram2.log
