#!/bin/bash
# K1 phase-1 prefetch forms: default (shared-memory-limited launches only), + direct launches, none
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/stage_detail.py > gpurun_out/sd_s5b_$name.txt 2>&1; echo == $name; grep -E " basis " gpurun_out/sd_s5b_$name.txt | awk '{printf "%s/%s/%s ", $3, $5, $6} END {print ""}'; grep -E "^basis|^\{" gpurun_out/sd_s5b_$name.txt; }
run default X=1
run pfdirect FOKL_BASIS_PF_DIRECT=1
run nopf FOKL_BASIS_NOPF=1
( timeout 300 python -m pytest tests/test_gpu_basis.py -m gpu -x -q 2>&1 | tail -3 )
( FOKL_BASIS_PF_DIRECT=1 timeout 300 python -m pytest tests/test_gpu_basis.py -m gpu -x -q 2>&1 | tail -3 )
