#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int8.py tests/test_gpu_api.py -q > gpurun_out/c26_tests.log 2>&1; tail -4 gpurun_out/c26_tests.log
REFTESTS="test_autograd test_linear8bitlt" timeout 900 bash tools/run_reference_tests.sh
