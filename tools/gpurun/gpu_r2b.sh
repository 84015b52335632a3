#!/bin/bash
# round 2: blocked eigensolver (eigbig.cuh) first contact
mkdir -p gpurun_out
( FOKL_EIGB_MIN_P=2 timeout 120 python tools/eig_big_check.py 17,64,130,300,500 1 2>&1 | grep -v Warn ) > gpurun_out/r2b_small.log; cat gpurun_out/r2b_small.log
( timeout 300 python tools/eig_big_check.py 705,1024,1700,2072 12 2>&1 | grep -v Warn ) > gpurun_out/r2b_big.log; cat gpurun_out/r2b_big.log
