#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 50 --warmup 5 > gpurun_out/c12_bench.log 2> gpurun_out/c12_bench.err; echo "bench rc=$?"
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c12_tests.log 2>&1; tail -3 gpurun_out/c12_tests.log
