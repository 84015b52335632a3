#!/bin/bash
# session 4, call j: kill loop from the eigendecomposition; bounce-buffer fix; full GPU suite + default bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r4j_pytest.log; cat gpurun_out/r4j_pytest.log
( timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r4j_bench.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4j_bench.log | head -3; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4j_bench.log; grep -o '"pageable": {[^}]*}' gpurun_out/r4j_bench.log
( FOKL_B200_KILL_FROM_EIG=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4j_bench_noeig.log
echo "kill_from_eig=0"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4j_bench_noeig.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4j_bench_noeig.log
( timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4j_bench_cfg5.log
echo cfg5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4j_bench_cfg5.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4j_bench_cfg5.log
