#!/bin/bash
# third GPU call: decoupled activation ring, XLOCAL A/B with traces, staged outlier epilogue, new tests
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/c3_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/c3_$name.log | cut -c1-160))"; }
step pair_trace  300 python tools/probe_gemm4_pair.py trace
step pair_time   500 python tools/probe_gemm4_pair.py time 4096x4096x4096 4096x11008x4096 4096x4096x11008 2048x14336x4096 1024x4096x4096 8192x8192x8192
step pair_eq     500 python tools/probe_gemm4_pair.py eq
step tests_new   900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_reference_loader.py -q
step bench_int8  400 python bench.py --workload int8_c3 --steps 30 --no-cpu-baseline
step reftests    2400 bash tools/run_reference_tests.sh
for f in gpurun_out/c3_pair_time.log gpurun_out/c3_tests_new.log gpurun_out/c3_bench_int8.log gpurun_out/reftests_summary.txt; do echo "---- $f"; tail -12 "$f" | cut -c1-900; done
grep -E "=====|MMA stage period|MMA thread|epilogue begin|decode:" gpurun_out/c3_pair_trace.log
grep -E "eq done|MISMATCH|UNEXPECTED" gpurun_out/c3_pair_eq.log | head
