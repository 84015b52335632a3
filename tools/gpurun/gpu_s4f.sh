#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python tools/stage_detail.py > gpurun_out/sd_s4f_$name.txt 2>&1; echo == $name; grep -E " gram " gpurun_out/sd_s4f_$name.txt | awk '{printf "%s/%s ", $3, $5} END {print ""}'; grep -E "^gram" gpurun_out/sd_s4f_$name.txt; }
run cpasync FOKL_GRAM_CPASYNC=1
run mb FOKL_X=1
run mb_cap128 FOKL_GRAM_CAP=128
run mb_cap96 FOKL_GRAM_CAP=96
run cpasync_cap128 FOKL_GRAM_CPASYNC=1 FOKL_GRAM_CAP=128
