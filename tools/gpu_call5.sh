#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/c5_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/c5_$name.log | cut -c1-160))"; }
step pair_eq     500 python tools/probe_gemm4_pair.py eq
step pair_trace  300 python tools/probe_gemm4_pair.py trace
step pair_time   500 python tools/probe_gemm4_pair.py time 4096x4096x4096 4096x11008x4096 4096x4096x11008 1024x4096x4096 8192x8192x8192
step tests_new   900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_reference_loader.py -q
for f in gpurun_out/c5_pair_time.log gpurun_out/c5_tests_new.log; do echo "---- $f"; tail -8 "$f" | cut -c1-1100; done
grep -E "MMA stage period|MMA thread|issue deltas|epilogue begin|decode:" gpurun_out/c5_pair_trace.log
grep -E "eq done|MISMATCH|UNEXPECTED" gpurun_out/c5_pair_eq.log | head
