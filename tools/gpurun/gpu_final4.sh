#!/bin/bash
# last verification of the round: smoke, full gpu suite, default bench line
mkdir -p gpurun_out
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/final4_smoke.log; cat gpurun_out/final4_smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/final4_pytest.log; cat gpurun_out/final4_pytest.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/final4_bench.log; cut -c1-300 gpurun_out/final4_bench.log
