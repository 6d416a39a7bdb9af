#!/bin/bash
mkdir -p gpurun_out
BNB_B200_PAIR_MT=256 ncu --set full --clock-control none --import-source on -k regex:gemm4_pair -s 2 -c 1 -o gpurun_out/r02_gemm4_pair_mt256 python tools/run_gemm4_one.py 4096 4096 4096 > gpurun_out/c10_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c10_ncu.log
python tools/run_i8_one.py
