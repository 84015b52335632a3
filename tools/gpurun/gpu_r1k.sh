#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_r1k.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_r1k.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1k.csv $B > gpurun_out/launches_r1k.log 2>&1
cat gpurun_out/pytest_r1k.log; for f in gpurun_out/bench_r1k.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; grep -o '"roofline": {[^}]*}' $f | cut -c1-200; done
