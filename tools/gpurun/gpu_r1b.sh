#!/bin/bash
# round-1 evidence pass: gpu tests, smoke, full bench line (+ reference arm), launch list, --set full of K1 / K2
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
cp /root/repo/MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 120 tools/micro/dmma_bench 2>&1 ) > gpurun_out/dmma_bench.log
( timeout 900 python bench.py 2>&1 | tail -3 ) > gpurun_out/bench_default.log
( timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -2 ) > gpurun_out/bench_reference.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches_bench.log 2>&1
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_kernel -c 13 -f -o gpurun_out/prof_basis $B0 > gpurun_out/prof_basis.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 8 -c 6 -f -o gpurun_out/prof_gram $B0 > gpurun_out/prof_gram.log 2>&1
tail -c 400 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/dmma_bench.log | tail -60; cat gpurun_out/bench_default.log
ls -la gpurun_out
