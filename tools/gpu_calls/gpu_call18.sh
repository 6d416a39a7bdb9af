#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm4.py tests/test_gpu_parity_full.py -q -x > gpurun_out/c18_tests.log 2>&1; tail -3 gpurun_out/c18_tests.log
timeout 400 python tools/probe_decode2.py > gpurun_out/c18_decode.log 2>&1; echo "rc=$?"; grep -h "^M\|WARPS" gpurun_out/c18_decode.log
