#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/kill_big_check.py 150,300,900,1500 1000 2>&1 | grep -v Warn ) > gpurun_out/r2e_kill.log; cat gpurun_out/r2e_kill.log
( timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 --cprofile 1 2>&1 | grep -v "^substage\|^(" | head -70 ) > gpurun_out/r2e_cfg5_prof.log; cat gpurun_out/r2e_cfg5_prof.log
