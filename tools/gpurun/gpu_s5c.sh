#!/bin/bash
# ncu --set full of the main-effect (C = 8, direct) and two-way (C = 28) basis launches
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:basis_kernel -s 0 -c 2 -f -o gpurun_out/prof_basis_s5c python tools/stage_detail.py > gpurun_out/prof_basis_s5c.log 2>&1
tail -3 gpurun_out/prof_basis_s5c.log; ls -la gpurun_out/prof_basis_s5c.ncu-rep
