#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5g_$name.txt 2>&1; echo == $name; tail -7 gpurun_out/sd_s5g_$name.txt; }
run pipe X=1
( timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 )
