#!/bin/bash
mkdir -p gpurun_out
for w in 8 16; do BNB_B200_GEMV_WARPS=$w timeout 400 python tools/probe_decode2.py > gpurun_out/c16_decode_w$w.log 2>&1; echo "w=$w rc=$?"; done
timeout 300 python -m pytest tests/test_gpu_gemm4.py -q -k "gemv or decode or mma or small" > gpurun_out/c16_tests.log 2>&1; tail -2 gpurun_out/c16_tests.log
grep -h "M\|WARPS" gpurun_out/c16_decode_w8.log gpurun_out/c16_decode_w16.log
