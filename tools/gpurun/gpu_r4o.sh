#!/bin/bash
# session 4, call o: one read-back per candidate batch; full GPU suite + bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r4o_pytest.log; cat gpurun_out/r4o_pytest.log
( timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4o_bench_cfg4.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4o_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4o_bench_cfg4.log
timeout 300 python tools/host_profile.py --top 30 > gpurun_out/r4o_host_profile.txt 2>&1; grep -A9 "main-stream stages busy" gpurun_out/r4o_host_profile.txt
