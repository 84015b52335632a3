#!/bin/bash
# session 4, call m: wide K2 tiles with deeper slabs / fewer ring stages
mkdir -p gpurun_out
timeout 900 python tools/gram_sweep.py --reps 5 --shapes '28,168;51,168;58,168' --variants 'auto;FOKL_GRAM_KB=32;FOKL_GRAM_KB=32,FOKL_GRAM_STAGES=3;FOKL_GRAM_STAGES=6;FOKL_GRAM_STAGES=5;FOKL_GRAM_STAGES=4' > gpurun_out/r4m_sweep.txt 2>&1
cat gpurun_out/r4m_sweep.txt
