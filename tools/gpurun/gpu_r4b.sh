#!/bin/bash
# session 4, call b: K2 loose blocks one per warp (FLEX) + sub-partition-balanced placement (FOKL_GRAM_PLACE=2)
mkdir -p gpurun_out
FOKL_GRAM_PLACE=2 timeout 900 python -m pytest tests/test_gpu_gram.py -x -q 2>&1 | tail -4 > gpurun_out/r4b_pytest_gram.log
cat gpurun_out/r4b_pytest_gram.log
timeout 900 python tools/gram_sweep.py --reps 6 --variants 'auto;FOKL_GRAM_PLACE=2;FOKL_GRAM_PLACE=2,FOKL_GRAM_WARPS=15;FOKL_GRAM_PLACE=2,FOKL_GRAM_WARPS=12;FOKL_GRAM_PLACE=1,FOKL_GRAM_WARPS=15' > gpurun_out/r4b_sweep.txt 2>&1
cat gpurun_out/r4b_sweep.txt
