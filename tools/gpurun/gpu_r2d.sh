#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/kill_big_check.py 150,300,900,2072 1000 2>&1 | grep -v Warn ) > gpurun_out/r2c_kill.log; cat gpurun_out/r2c_kill.log
( timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 2>&1 | tail -25 ) > gpurun_out/r2c_cfg5.log; cat gpurun_out/r2c_cfg5.log
( timeout 300 python tools/explore.py --cfg cfg5 --n 1000000 --resident 1 2>&1 | tail -25 ) > gpurun_out/r2c_cfg5_1m.log; cat gpurun_out/r2c_cfg5_1m.log
