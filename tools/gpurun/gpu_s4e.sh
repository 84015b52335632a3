#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel_mb -s 3 -c 4 -f -o gpurun_out/prof_gram_s4e $B > gpurun_out/prof_gram_s4e.log 2>&1
tail -3 gpurun_out/prof_gram_s4e.log
