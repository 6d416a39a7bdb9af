#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_optim.py -q -s -p no:cacheprovider > gpurun_out/c21_optim.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c21_optim.log
grep -E "^FAILED|Error|identical" gpurun_out/c21_optim.log | cut -c1-220 | sort | uniq -c | sort -rn | head -60
REFTESTS=test_optim timeout 1500 bash tools/run_reference_tests.sh; tail -30 gpurun_out/reftests_test_optim.log | cut -c1-200
