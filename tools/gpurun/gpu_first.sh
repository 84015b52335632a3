#!/bin/bash
# first GPU pass: tests, smoke, per-stage timing of cfg4, first bench line
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -20 ) > gpurun_out/smoke.log
( timeout 600 python tools/explore.py --cfg cfg4 --n 1000000 2>&1 | tail -80 ) > gpurun_out/explore_1m.log
( timeout 900 python tools/explore.py --cfg cfg4 2>&1 | tail -80 ) > gpurun_out/explore_10m.log
( timeout 900 python bench.py --steps 1 --warmup 1 2>&1 | tail -20 ) > gpurun_out/bench_first.log
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_first.log
