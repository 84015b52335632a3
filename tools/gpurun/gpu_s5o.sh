#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5o.txt 2>&1; tail -7 gpurun_out/sd_s5o.txt
grep -E "candidates_chain|kill_loop|compact" gpurun_out/sd_s5o.txt | awk '{printf "%s %s %s | ", $1, $2, $3} NR%3==0 {print ""}' | tail -8
( timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_candidates.py -m gpu -x -q 2>&1 | tail -3 )
