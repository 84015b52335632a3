#!/bin/bash
# round 2, first call: state of the tree on a fresh box + where cfg5 spends its time today
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
( timeout 300 python tools/explore.py --cfg cfg5 --n 200000 --resident 1 2>&1 | tail -40 ) > gpurun_out/r2a_cfg5.log; cat gpurun_out/r2a_cfg5.log
