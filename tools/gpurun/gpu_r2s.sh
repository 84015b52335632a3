#!/bin/bash
mkdir -p gpurun_out
( FOKL_B200_DEBUG=1 timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 2>&1 | grep -v "^substage\|^(" | tail -14 ) > gpurun_out/r2s_cfg5.log; cat gpurun_out/r2s_cfg5.log
