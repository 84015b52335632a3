#!/bin/bash
# ncu evidence: launch list of one bench step (cfg4, 10M rows) + full-set captures of the hot kernels (2M rows)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
for spec in "basis_kernel:6:2:basis" "gram_kernel:7:1:gram" "cand_eig_kernel:12:1:eig" "cand_chain_kernel:12:1:chain" "kill_scores_kernel:300:1:kill" "cand_betas_kernel:12:1:betas"; do
  IFS=: read k s c name <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f -o gpurun_out/prof_$name $B --n 2000000 > gpurun_out/prof_$name.log 2>&1
done
ls -la gpurun_out
