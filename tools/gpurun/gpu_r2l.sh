#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fit.py tests/test_gpu_fullsize.py -m gpu -x -q --durations=5 2>&1 | tail -25 ) > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
for pf in 1 0; do
( FOKL_B200_PREFETCH=$pf timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2l_bench_pf$pf.log
echo "prefetch=$pf"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2l_bench_pf$pf.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2l_bench_pf$pf.log
done
