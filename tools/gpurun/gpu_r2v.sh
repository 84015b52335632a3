#!/bin/bash
mkdir -p gpurun_out
for mp in 1000 190 120; do
( FOKL_EIGB_MIN_P=$mp timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2v_bench_$mp.log
echo "FOKL_EIGB_MIN_P=$mp"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2v_bench_$mp.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2v_bench_$mp.log
done
( timeout 100 python tools/eig_diag.py 100,128,160,200,226 2>&1 | grep -v Warn ) > gpurun_out/r2v_eigdiag_cluster.log; cat gpurun_out/r2v_eigdiag_cluster.log
( FOKL_EIGB_MIN_P=2 timeout 100 python tools/eig_diag.py 100,128,160,200,226 2>&1 | grep -v Warn ) > gpurun_out/r2v_eigdiag_eigb.log; cat gpurun_out/r2v_eigdiag_eigb.log
