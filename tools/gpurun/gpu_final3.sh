#!/bin/bash
# round-1 final evidence (session 5): bench lines, launch list, DRAM traffic of every K1 / K2 launch, --set full captures
mkdir -p gpurun_out
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/final3_bench.log
( timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -1 ) > gpurun_out/final3_bench_reference.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final3_launches.csv $B > gpurun_out/final3_launches.log 2>&1
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"basis_kernel|gram_kernel" --csv --log-file gpurun_out/final3_traffic.csv $B0 > gpurun_out/final3_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:basis_kernel -s 6 -c 1 -f -o gpurun_out/final3_prof_basis $B0 > gpurun_out/final3_prof_basis.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 11 -c 1 -f -o gpurun_out/final3_prof_gram $B0 > gpurun_out/final3_prof_gram.log 2>&1
ls -la gpurun_out | grep final3; cut -c1-600 gpurun_out/final3_bench.log; cut -c1-300 gpurun_out/final3_bench_reference.log
