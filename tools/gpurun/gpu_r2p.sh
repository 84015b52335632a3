#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fit.py -m gpu -x -q -k "nested or build_ahead" 2>&1 | tail -8 ) > gpurun_out/r2p_pytest.log; cat gpurun_out/r2p_pytest.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2p_bench_cfg4.log
echo cfg4; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2p_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2p_bench_cfg4.log
