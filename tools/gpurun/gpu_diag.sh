#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python tools/eig_diag.py 2>&1 | tail -40 ) > gpurun_out/eig_diag.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cand_eigj_kernel -c 1 -f -o gpurun_out/prof_eigj python tools/eig_diag.py 230 > gpurun_out/prof_eigj.log 2>&1
cat gpurun_out/eig_diag.log
