#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_s5d.log
cat gpurun_out/pytest_s5d.log
