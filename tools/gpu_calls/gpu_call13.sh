#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/c13_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/c13_$name.log | cut -c1-160))"; }
step pair_eq     500 python tools/probe_gemm4_pair.py eq
step lite        300 python tools/probe_gemm4_pair.py lite
step pair_time   500 python tools/probe_gemm4_pair.py time 4096x4096x4096 4096x11008x4096 4096x4096x11008 1024x4096x4096 8192x8192x8192
step tests_new   900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_gemm4.py -q
for f in gpurun_out/c13_pair_time.log gpurun_out/c13_tests_new.log; do echo "---- $f"; tail -7 "$f" | cut -c1-1100; done
grep -E "lite|timeline|round|gap" gpurun_out/c13_lite.log
grep -E "eq done|MISMATCH|UNEXPECTED" gpurun_out/c13_pair_eq.log | head -4
