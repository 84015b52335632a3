#!/bin/bash
# round-1 evidence, profiler part: launch list, DRAM traffic of every K1 / K2 launch, --set full of the dominant launches
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/final_pytest_fit.log
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/final_bench_b.log
( timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 ) > gpurun_out/final_bench_reference.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $B > gpurun_out/final_launches.log 2>&1
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"basis_kernel|gram_kernel" --csv --log-file gpurun_out/final_traffic.csv $B0 > gpurun_out/final_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:basis_kernel -s 6 -c 1 -f -o gpurun_out/final_prof_basis $B0 > gpurun_out/final_prof_basis.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 11 -c 1 -f -o gpurun_out/final_prof_gram $B0 > gpurun_out/final_prof_gram.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cand_chain_warp -s 12 -c 1 -f -o gpurun_out/final_prof_chain $B0 > gpurun_out/final_prof_chain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cand_eigj -s 12 -c 2 -f -o gpurun_out/final_prof_eigj $B0 > gpurun_out/final_prof_eigj.log 2>&1
ls -la gpurun_out; cat gpurun_out/final_pytest_fit.log; cat gpurun_out/final_bench_b.log | cut -c1-400
