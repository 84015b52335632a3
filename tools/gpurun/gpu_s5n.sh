#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_s5n.log; cat gpurun_out/pytest_s5n.log
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5n.txt 2>&1; tail -7 gpurun_out/sd_s5n.txt
