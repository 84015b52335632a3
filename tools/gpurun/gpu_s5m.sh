#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/eig_diag.py 30,100,220,440 2>&1 | grep -v chain | tail -5 )
( timeout 600 python -m pytest tests/test_gpu_candidates.py tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -3 )
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5m.txt 2>&1; tail -7 gpurun_out/sd_s5m.txt | head -5
