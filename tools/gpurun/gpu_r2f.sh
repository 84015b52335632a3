#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 2>&1 | grep -v "^substage\|^(" | tail -12 ) > gpurun_out/r2f_cfg5.log; cat gpurun_out/r2f_cfg5.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eigb_kernel -s 1 -c 1 -o gpurun_out/r2f_eigb python tools/eig_big_check.py 2072 1 gauss > gpurun_out/r2f_ncu.log 2>&1; tail -5 gpurun_out/r2f_ncu.log
