#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/smoke_s5j.log; cat gpurun_out/smoke_s5j.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/pytest_s5j.log; cat gpurun_out/pytest_s5j.log
