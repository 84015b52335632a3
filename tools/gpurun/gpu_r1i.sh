#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_r1i.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_r1i.log
( timeout 400 python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -3 ) > gpurun_out/bench_cfg5.log
( timeout 300 python bench.py --workload cfg3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -2 ) > gpurun_out/bench_cfg3.log
cat gpurun_out/pytest_r1i.log; for f in gpurun_out/bench_r1i.log gpurun_out/bench_cfg5.log gpurun_out/bench_cfg3.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f;  grep -o '"candidate_models_per_step": [0-9.]*, "terms_selected": [0-9]*, "substages": [0-9]*' $f; tail -c 300 $f; done
