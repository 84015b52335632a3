#!/bin/bash
# round-1 final evidence: gpu tests, smoke, bench line (+ cpu_baseline), reference arm, launch list, --set full of K1 / K2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/final_gpu.txt 2>&1
ls -la /root/repo/MEASURED_PEAKS.json >> gpurun_out/final_gpu.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > gpurun_out/final_pytest.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/final_smoke.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/final_bench.log
( timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 ) > gpurun_out/final_bench_reference.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/final_launches.log 2>&1
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:basis_kernel -c 13 -f -o gpurun_out/final_prof_basis $B0 > gpurun_out/final_prof_basis.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -c 14 -f -o gpurun_out/final_prof_gram $B0 > gpurun_out/final_prof_gram.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cand_chain_warp -s 12 -c 2 -f -o gpurun_out/final_prof_chain $B0 > gpurun_out/final_prof_chain.log 2>&1
cat gpurun_out/final_pytest.log gpurun_out/final_smoke.log gpurun_out/final_bench.log
