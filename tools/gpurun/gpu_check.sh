#!/bin/bash
# tests + short bench + host profile + diagnostics
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log
( timeout 600 python tools/eig_diag.py 2>&1 | tail -40 ) > gpurun_out/eig_diag.log
( timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -5 ) > gpurun_out/bench_check.log
( timeout 600 python tools/explore.py --cfg cfg4 --repeat 2 --cprofile 1 2>&1 | tail -90 ) > gpurun_out/explore_10m.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
tail -c 600 gpurun_out/pytest_gpu.log; cat gpurun_out/eig_diag.log; tail -c 3500 gpurun_out/bench_check.log
