#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python tools/eig_batch_diag.py 2>&1 | tail -30 ) > gpurun_out/eig_batch_diag.log
( timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_candidates.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_r1e.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_r1e.log
( timeout 600 python tools/explore.py --cfg cfg4 --repeat 3 --resident 1 --cprofile 1 2>&1 | tail -70 ) > gpurun_out/explore_resident2.log
cat gpurun_out/eig_batch_diag.log gpurun_out/pytest_r1e.log gpurun_out/bench_r1e.log; head -30 gpurun_out/explore_resident2.log
