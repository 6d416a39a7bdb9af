#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_optim.py tests/test_gpu_blockwise.py -q > gpurun_out/c24_tests.log 2>&1; tail -3 gpurun_out/c24_tests.log
timeout 300 python bench.py --workload optim_f4 > gpurun_out/c24_optim_bench.log 2> gpurun_out/c24_optim_bench.err; echo "optim rc=$?"
