#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2m_bench.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2m_bench.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2m_bench.log
( timeout 300 python tools/explore.py --cfg cfg4 --resident 1 --repeat 3 --cprofile 1 2>&1 | grep -v "^substage\|^(" | head -75 ) > gpurun_out/r2m_cfg4_prof.log; cat gpurun_out/r2m_cfg4_prof.log
