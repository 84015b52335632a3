#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/stage_detail.py > gpurun_out/stage_detail_s4b.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s4b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_s4b.log 2>&1
tail -5 gpurun_out/ncu_s4b.log
cat gpurun_out/stage_detail_s4b.txt | tail -80
