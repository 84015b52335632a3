#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gram.py tests/test_gpu_candidates.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_s4d.log
cat gpurun_out/pytest_s4d.log
timeout 300 python tools/stage_detail.py > gpurun_out/stage_detail_s4d_mb.txt 2>&1
FOKL_GRAM_CPASYNC=1 timeout 300 python tools/stage_detail.py > gpurun_out/stage_detail_s4d_cpasync.txt 2>&1
for f in gpurun_out/stage_detail_s4d_mb.txt gpurun_out/stage_detail_s4d_cpasync.txt; do echo == $f; grep -E " gram " $f; tail -7 $f; done
