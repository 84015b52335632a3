#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_s4p.log
cat gpurun_out/pytest_s4p.log
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s4p.txt 2>&1; grep -E " gram " gpurun_out/sd_s4p.txt | awk '{printf "%s/%s ", $3, $5} END {print ""}'; tail -7 gpurun_out/sd_s4p.txt
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel_tma -s 1 -c 1 -f -o gpurun_out/prof_gram_s4p $B > gpurun_out/prof_gram_s4p.log 2>&1
tail -2 gpurun_out/prof_gram_s4p.log
