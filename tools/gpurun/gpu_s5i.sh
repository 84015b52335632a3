#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/micro/dmma_dfma_mix > gpurun_out/dmma_dfma_mix.txt 2>&1; cat gpurun_out/dmma_dfma_mix.txt
( timeout 300 python tools/eig_diag.py 100,220 2>&1 | tail -6 )
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cand_eigj_kernel -s 1 -c 1 -f -o gpurun_out/prof_eigj_s5i python tools/eig_diag.py 220 > gpurun_out/prof_eigj_s5i.log 2>&1
tail -3 gpurun_out/prof_eigj_s5i.log
