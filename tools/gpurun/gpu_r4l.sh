#!/bin/bash
# session 4, call l: full model vs side batch contention -- high-priority main candidate stream, SMs reserved from the side batch
mkdir -p gpurun_out
run() { ( env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4l_tmp.log; echo "$@"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4l_tmp.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4l_tmp.log; }
run FOKL_B200_MAIN_HP=0
run FOKL_B200_MAIN_HP=1
run FOKL_B200_SIDE_RESERVE=32
run FOKL_B200_SIDE_RESERVE=0
run FOKL_B200_MAIN_HP=1 FOKL_B200_SIDE_RESERVE=32
