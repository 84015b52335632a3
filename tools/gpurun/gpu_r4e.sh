#!/bin/bash
# session 4, call e: ncu --set full of the (28 old, 168 new) K2 launch after the placement changes
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel_tma -s 1 -c 1 -f -o gpurun_out/r4e_prof_gram python tools/gram_sweep.py --reps 1 --shapes '28,168' --variants auto > gpurun_out/r4e_prof_gram.log 2>&1
tail -3 gpurun_out/r4e_prof_gram.log
ls -la gpurun_out/r4e*
