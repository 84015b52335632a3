#!/bin/bash
mkdir -p gpurun_out
( FOKL_EIGB_MIN_P=2 timeout 120 python tools/eig_big_check.py 17,64,130,300 1 gauss 2>&1 | grep -v Warn ) > gpurun_out/r2g_small.log; cat gpurun_out/r2g_small.log
( timeout 300 python tools/eig_big_check.py 705,1024,2072 12 gauss,spline 2>&1 | grep -v Warn ) > gpurun_out/r2g_big.log; cat gpurun_out/r2g_big.log
( timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 2>&1 | grep -v "^substage\|^(" | tail -9 ) > gpurun_out/r2g_cfg5.log; cat gpurun_out/r2g_cfg5.log
