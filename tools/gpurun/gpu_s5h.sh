#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5h.txt 2>&1; tail -7 gpurun_out/sd_s5h.txt
timeout 300 python tools/explore.py --cfg cfg4 --repeat 3 --cprofile 1 --resident 1 > gpurun_out/explore_s5h.txt 2>&1
grep -v Warn gpurun_out/explore_s5h.txt | sed -n 1,70p | cut -c1-150
