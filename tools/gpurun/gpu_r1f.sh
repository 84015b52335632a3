#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_candidates.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_r1f.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_r1f.log
( timeout 600 python tools/explore.py --cfg cfg4 --repeat 3 --cprofile 1 2>&1 | tail -70 ) > gpurun_out/explore_e2e.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv $B > gpurun_out/launches_r1f.log 2>&1
cat gpurun_out/pytest_r1f.log gpurun_out/bench_r1f.log; head -45 gpurun_out/explore_e2e.log
