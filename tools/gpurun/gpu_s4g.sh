#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gram_sweep.py --variants 'auto;FOKL_GRAM_KERNEL=cpasync;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=64;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=128;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=128,FOKL_GRAM_STAGES=2;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=256,FOKL_GRAM_STAGES=2' > gpurun_out/gram_sweep_s4g.txt 2>&1
cat gpurun_out/gram_sweep_s4g.txt
