#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm4_pair -s 2 -c 1 -o gpurun_out/r02_gemm4_pair_persistent -f python tools/run_gemm4_one.py 4096 4096 4096 > gpurun_out/c17_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/c17_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-secondary > gpurun_out/c17_launches.log 2>&1; echo "launch list rc=$?"; tail -c 300 gpurun_out/c17_launches.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_int8_fwd.csv python tools/run_i8_fwd.py > gpurun_out/c17_i8.log 2>&1; echo "i8 rc=$?"; tail -3 gpurun_out/c17_i8.log
