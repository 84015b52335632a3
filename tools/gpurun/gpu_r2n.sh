#!/bin/bash
mkdir -p gpurun_out
( for a in "60 10 gauss" "400 30 gauss" "400 30 spline" "1200 40 spline" "2072 60 gauss"; do timeout 300 python tools/nested_check.py $a 2>&1 | grep -v Warn; done ) > gpurun_out/r2n_nested.log; cat gpurun_out/r2n_nested.log
