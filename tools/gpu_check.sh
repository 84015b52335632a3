#!/bin/bash
# tests + short bench
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 ) > gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -5 ) > gpurun_out/bench_check.log
( timeout 600 python tools/explore.py --cfg cfg4 2>&1 | tail -40 ) > gpurun_out/explore_10m.log
tail -c 1500 gpurun_out/pytest_gpu.log; tail -c 3000 gpurun_out/bench_check.log
