#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 8 -f -o gpurun_out/prof_gram_s4c $B > gpurun_out/prof_gram_s4c.log 2>&1
tail -3 gpurun_out/prof_gram_s4c.log; ls -la gpurun_out/*.ncu-rep
