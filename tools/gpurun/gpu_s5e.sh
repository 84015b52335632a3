#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/pytest_s5e.log
tail -30 gpurun_out/pytest_s5e.log
