#!/bin/bash
# second GPU call of round 2: the reworked pair kernel (direct relaxed arrives, TMA-store epilogue, 2-way split)
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/c2_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/c2_$name.log | cut -c1-160))"; }
step pair_eq     500 python tools/probe_gemm4_pair.py eq
step pair_trace  200 python tools/probe_gemm4_pair.py trace
step pair_time   400 python tools/probe_gemm4_pair.py time
step tests_new   900 python -m pytest tests/test_gpu_parity_full.py -q -x
step tests       900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity_full.py
step bench       900 python bench.py --steps 50 --warmup 5
for f in gpurun_out/c2_pair_eq.log gpurun_out/c2_pair_time.log gpurun_out/c2_tests_new.log gpurun_out/c2_tests.log; do echo "---- $f"; tail -12 "$f" | cut -c1-600; done
grep -E "MMA stage period|epilogue begin|decode:" gpurun_out/c2_pair_trace.log
