#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm4_pair -s 2 -c 1 -o gpurun_out/r02_gemm4_pair_m4096 python tools/run_gemm4_one.py 4096 4096 4096 > gpurun_out/c9_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c9_ncu.log
ncu --set full --clock-control none -k regex:int8_gemm -s 2 -c 1 -o gpurun_out/r02_int8_gemm python tools/run_i8_one.py > gpurun_out/c9_ncu_i8.log 2>&1
echo "ncu i8 rc=$?"; tail -3 gpurun_out/c9_ncu_i8.log
timeout 600 python bench.py --workload int8_c3 --steps 30 --no-cpu-baseline > gpurun_out/c9_bench_int8.log 2>&1; tail -c 700 gpurun_out/c9_bench_int8.log
