#!/bin/bash
# session 4, call q: last verification chains enqueued before the wait for the previous batch
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fullsize.py tests/test_gpu_update.py -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r4q_pytest.log; cat gpurun_out/r4q_pytest.log
( timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4q_bench_cfg4.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4q_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4q_bench_cfg4.log
( timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4q_bench_cfg5.log
echo cfg5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4q_bench_cfg5.log | head -1
timeout 300 python tools/host_profile.py --top 30 > gpurun_out/r4q_host_profile.txt 2>&1; grep -A9 "main-stream stages busy" gpurun_out/r4q_host_profile.txt
