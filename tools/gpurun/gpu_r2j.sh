#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -14 ) > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
