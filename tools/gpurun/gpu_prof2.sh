#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gram.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_gram.log
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
for spec in "basis_kernel:6:1:basis" "cand_eigj_kernel:12:1:eigj" "cand_chain_kernel:12:1:chain"; do
  IFS=: read k s c name <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_$name $B > gpurun_out/prof_$name.log 2>&1
done
cat gpurun_out/pytest_gram.log
