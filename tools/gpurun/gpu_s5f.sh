#!/bin/bash
# pipelined selection loop: full gpu suite, then one cfg4 fit per form
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/pytest_s5f.log
tail -12 gpurun_out/pytest_s5f.log
run() { name=$1; shift; env "$@" timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5f_$name.txt 2>&1; echo == $name; tail -7 gpurun_out/sd_s5f_$name.txt; }
run pipe X=1
run seq FOKL_B200_PIPELINE=0
