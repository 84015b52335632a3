#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4 )
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5k.txt 2>&1; grep -E " gram " gpurun_out/sd_s5k.txt | awk '{printf "%s/%s ", $3, $5} END {print ""}'; tail -7 gpurun_out/sd_s5k.txt
