#!/bin/bash
# round-2 evidence: smoke, gpu suite, bench lines (cfg4 both arms, cfg5), launch list, DRAM traffic of every K1 / K2
# launch (with the commit hash), one --set full capture of the Gram kernel
mkdir -p gpurun_out
cp tools/.commit gpurun_out/f2_commit.txt 2>/dev/null
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/f2_smoke.log; cat gpurun_out/f2_smoke.log
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > gpurun_out/f2_pytest.log; cat gpurun_out/f2_pytest.log
( timeout 900 python bench.py 2>&1 | tail -1 ) > gpurun_out/f2_bench.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 ) > gpurun_out/f2_bench_reference.log
( timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/f2_bench_cfg5.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f2_launches.csv $B > gpurun_out/f2_launches.log 2>&1
B0="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"basis_kernel|gram_kernel" --csv --log-file gpurun_out/f2_traffic.csv $B0 > gpurun_out/f2_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 11 -c 1 -f -o gpurun_out/f2_prof_gram $B0 > gpurun_out/f2_prof_gram.log 2>&1
ls -la gpurun_out | grep f2_; cut -c1-400 gpurun_out/f2_bench.log; cut -c1-300 gpurun_out/f2_bench_reference.log; cut -c1-300 gpurun_out/f2_bench_cfg5.log
