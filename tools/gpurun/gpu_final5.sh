#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final5_pytest.log; cat gpurun_out/final5_pytest.log
