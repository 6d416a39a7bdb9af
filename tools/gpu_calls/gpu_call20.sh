#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int8.py tests/test_gpu_parity_full.py -q -x > gpurun_out/c20_tests.log 2>&1; tail -3 gpurun_out/c20_tests.log
timeout 200 python tools/run_i8_one.py > gpurun_out/c20_i8.log 2>&1; tail -1 gpurun_out/c20_i8.log
timeout 200 python tools/run_i8_one.py 4096 4096 4096 >> gpurun_out/c20_i8.log 2>&1; tail -1 gpurun_out/c20_i8.log
timeout 300 python bench.py --workload int8_c3 --steps 30 > gpurun_out/c20_int8_c3.log 2>&1; tail -c 900 gpurun_out/c20_int8_c3.log
