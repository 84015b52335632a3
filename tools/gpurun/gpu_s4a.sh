#!/bin/bash
# session-4 baseline: gpu tests + N=1 bench at the restored state
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_s4a.log
( timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 ) > gpurun_out/bench_s4a.log
cat gpurun_out/pytest_s4a.log; f=gpurun_out/bench_s4a.log; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; grep -o '"roofline": {[^}]*}' $f
