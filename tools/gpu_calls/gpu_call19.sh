#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:gemv4_mma -s 1 -c 1 -f -o gpurun_out/r02_gemv_mma_prmt python tools/run_decode_one.py 1 14336 4096 3 > gpurun_out/c19_a.log 2>&1; echo "prmt rc=$?"
BNB_B200_LIBRARY=$PWD/build_exp/libbnb_lut8.so ncu --set full --clock-control none -k regex:gemv4_mma -s 1 -c 1 -f -o gpurun_out/r02_gemv_mma_lut8 python tools/run_decode_one.py 1 14336 4096 3 > gpurun_out/c19_b.log 2>&1; echo "lut8 rc=$?"; tail -2 gpurun_out/c19_b.log
ncu --set full --clock-control none -k regex:gemv4_fast -s 1 -c 1 -f -o gpurun_out/r02_gemv_fast_m1 python tools/run_decode_one.py 1 14336 4096 0 > gpurun_out/c19_c.log 2>&1; echo "fast rc=$?"
