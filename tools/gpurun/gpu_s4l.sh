#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gram_sweep.py --variants "auto;FOKL_GRAM_PLACE=1" > gpurun_out/gram_sweep_s4l.txt 2>&1
cat gpurun_out/gram_sweep_s4l.txt
