#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_s4m.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_s4m.log
cat gpurun_out/pytest_s4m.log; f=gpurun_out/bench_s4m.log; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; grep -o '"roofline_gram": {[^}]*}' $f
