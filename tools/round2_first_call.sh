#!/bin/bash
# First GPU call of round 2 (one B200):  /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/r2_*.log.  Every new kernel uses bounded mbarrier waits (report + trap after
# 10 s) and each step has its own timeout.
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/r2_$name.log | cut -c1-160))"; }

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_smi.txt 2>&1
step pair_eq        600 python tools/probe_gemm4_pair.py eq
step pair_time      400 python tools/probe_gemm4_pair.py time
step pair_trace     200 python tools/probe_gemm4_pair.py trace
step tests          1200 python -m pytest tests -q -m gpu -x
step smoke          120 python -c "import __graft_entry__ as g; g.smoke()"
step bench          900 python bench.py --steps 50 --warmup 5
# 8-bit quantize with the cheaper (CPU-proven) search: the strict bit-exactness tests against the reference CUDA library
step q8fast_tests   300 env BNB_B200_Q8_FAST=1 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_zz_golden.py -q -k "8bit or None or quant"
step q8fast_bench   300 env BNB_B200_Q8_FAST=1 python bench.py --workload blockwise_c1 --no-cpu-baseline
for f in gpurun_out/r2_pair_eq.log gpurun_out/r2_pair_time.log gpurun_out/r2_pair_trace.log gpurun_out/r2_tests.log gpurun_out/r2_bench.log gpurun_out/r2_q8fast_bench.log; do echo "---- $f"; tail -15 "$f" | cut -c1-400; done
