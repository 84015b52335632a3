#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_s4r.log
cat gpurun_out/pytest_s4r.log
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s4r.txt 2>&1; grep -E "kill_loop|candidates" gpurun_out/sd_s4r.txt | awk '{printf "%s:%s ", $2, $3} END {print ""}'; tail -7 gpurun_out/sd_s4r.txt
