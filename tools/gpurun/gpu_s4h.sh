#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gram_sweep.py --shapes '28,168;51,168;58,168;37,56;64,56' --variants 'FOKL_GRAM_KERNEL=cpasync;FOKL_GRAM_KERNEL=cpasync,FOKL_GRAM_PLACE=1;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=32,FOKL_GRAM_STAGES=3;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=32,FOKL_GRAM_STAGES=3,FOKL_GRAM_PLACE=1;auto;FOKL_GRAM_PLACE=1' > gpurun_out/gram_sweep_s4h.txt 2>&1
cat gpurun_out/gram_sweep_s4h.txt
