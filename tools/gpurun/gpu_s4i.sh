#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/stage_detail.py > gpurun_out/sd_s4i_$name.txt 2>&1; echo == $name; grep -E " gram " gpurun_out/sd_s4i_$name.txt | awk '{printf "%s/%s ", $3, $5} END {print ""}'; tail -7 gpurun_out/sd_s4i_$name.txt; }
run auto FOKL_X=1
run place1 FOKL_GRAM_PLACE=1
