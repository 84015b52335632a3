#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_s4o.log
cat gpurun_out/pytest_s4o.log
timeout 600 python tools/gram_sweep.py --variants "auto;FOKL_GRAM_KERNEL=mb;FOKL_GRAM_KERNEL=cpasync;FOKL_GRAM_KB=16;FOKL_GRAM_KB=32;FOKL_GRAM_KB=64" > gpurun_out/gram_sweep_s4o.txt 2>&1
cat gpurun_out/gram_sweep_s4o.txt
