#!/bin/bash
mkdir -p gpurun_out
V='auto'
S='28,168;51,168;58,168;37,56'
FOKL_B200_LIB=tools/micro/_variants/libfokl_b200_old.so timeout 900 python tools/gram_sweep.py --shapes "$S" --variants "$V" > gpurun_out/gram_sweep_s4k_old.txt 2>&1
timeout 900 python tools/gram_sweep.py --shapes "$S" --variants "$V;FOKL_GRAM_PLACE=1" > gpurun_out/gram_sweep_s4k_new.txt 2>&1
FOKL_B200_LIB=tools/micro/_variants/libfokl_b200_old.so timeout 900 python tools/gram_sweep.py --shapes "$S" --variants "$V" > gpurun_out/gram_sweep_s4k_old2.txt 2>&1
cat gpurun_out/gram_sweep_s4k_old.txt gpurun_out/gram_sweep_s4k_new.txt gpurun_out/gram_sweep_s4k_old2.txt
