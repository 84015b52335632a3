#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_gram.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/pytest_s4n.log
cat gpurun_out/pytest_s4n.log
timeout 900 python tools/gram_sweep.py --variants "auto;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=32,FOKL_GRAM_STAGES=3;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=16,FOKL_GRAM_STAGES=4;FOKL_GRAM_KERNEL=mb,FOKL_GRAM_KB=64" > gpurun_out/gram_sweep_s4n.txt 2>&1
cat gpurun_out/gram_sweep_s4n.txt
