#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_basis.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final6_pytest.log; cat gpurun_out/final6_pytest.log
