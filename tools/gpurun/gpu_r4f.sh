#!/bin/bash
# session 4, call f: mid-size K2 shapes on the tensor-map kernel / other slab depths
mkdir -p gpurun_out
timeout 900 python tools/gram_sweep.py --reps 5 --shapes '8,28;14,56;20,56;37,28;37,56;64,56;49,8' --variants 'auto;FOKL_GRAM_KERNEL=tma;FOKL_GRAM_KERNEL=tma,FOKL_GRAM_KB=32;FOKL_GRAM_KERNEL=tma,FOKL_GRAM_KB=64;FOKL_GRAM_KB=128;FOKL_GRAM_KB=64' > gpurun_out/r4f_sweep.txt 2>&1
cat gpurun_out/r4f_sweep.txt
