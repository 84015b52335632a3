#!/bin/bash
# session 5: derivative kernel + K1 input prefetch -- full gpu suite, then the per-stage table of one cfg4 fit
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_s5a.log
tail -8 gpurun_out/pytest_s5a.log
timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5a.txt 2>&1; grep -E " basis " gpurun_out/sd_s5a.txt | awk '{printf "%s/%s/%s ", $3, $5, $6} END {print ""}'; tail -7 gpurun_out/sd_s5a.txt
