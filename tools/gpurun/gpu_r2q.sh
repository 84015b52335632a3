#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_candidates.py -m gpu -x -q -k "nested" 2>&1 | tail -12 ) > gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_pytest.log
