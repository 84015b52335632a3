#!/bin/bash
# round 2: whole-device kill loop vs single CTA; eigensolver on graded designs; inner-sweep knob; cfg5 explore
mkdir -p gpurun_out
( timeout 200 python tools/kill_big_check.py 150,300,900,2072 1000 2>&1 | grep -v Warn ) > gpurun_out/r2c_kill.log; cat gpurun_out/r2c_kill.log
( timeout 200 python tools/eig_big_check.py 1024,2072 1 spline 2>&1 | grep -v Warn ) > gpurun_out/r2c_eig_spline.log; cat gpurun_out/r2c_eig_spline.log
( FOKL_EIGB_INNER=2 timeout 200 python tools/eig_big_check.py 1024,2072 1 gauss,spline 2>&1 | grep -v Warn ) > gpurun_out/r2c_eig_inner2.log; cat gpurun_out/r2c_eig_inner2.log
( timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 2>&1 | tail -25 ) > gpurun_out/r2c_cfg5.log; cat gpurun_out/r2c_cfg5.log
