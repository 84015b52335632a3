#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_r1g.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_r1g.log
for v in nomask late both; do
  ( FOKL_B200_LIB=tools/micro/_variants/libfokl_b200_$v.so timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_r1g_$v.log
done
cat gpurun_out/pytest_r1g.log; for f in gpurun_out/bench_r1g*.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; done
