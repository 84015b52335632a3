#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2x_nested_launches.csv python tools/nested_check.py 2072 12 gauss > gpurun_out/r2x.log 2>&1
tail -2 gpurun_out/r2x.log
