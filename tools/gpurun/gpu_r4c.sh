#!/bin/bash
# session 4, call c: K2 k-split sweep under the new placement
mkdir -p gpurun_out
timeout 900 python tools/gram_sweep.py --reps 5 --shapes '1,8;8,28;12,8;14,56;20,56;21,8;37,28;37,56;49,8;64,56' --variants 'auto;FOKL_GRAM_KSPLIT=1;FOKL_GRAM_KSPLIT=2;FOKL_GRAM_KSPLIT=4;FOKL_GRAM_KSPLIT=8;FOKL_GRAM_KSPLIT=16' > gpurun_out/r4c_sweep.txt 2>&1
cat gpurun_out/r4c_sweep.txt
