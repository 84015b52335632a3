#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eigb_kernel -s 1 -c 1 -o gpurun_out/r2h_eigb python tools/eig_big_check.py 2072 1 gauss > gpurun_out/r2h_ncu.log 2>&1; tail -3 gpurun_out/r2h_ncu.log
