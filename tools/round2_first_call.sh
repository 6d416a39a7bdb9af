#!/bin/bash
# Everything that was written after round 1's GPU budget ran out, in one gpurun call (one B200):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Results land in gpurun_out/r2_*.log.  Nothing here can hang: every new kernel uses bounded mbarrier
# waits (report + trap after 10 s) and each step has its own timeout.
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/r2_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/r2_$name.log | cut -c1-160))"; }

step tests          900 python -m pytest tests -q -m gpu
step smoke          120 python -c "import __graft_entry__ as g; g.smoke()"
step bench          400 python bench.py
step bench_c1       300 python bench.py --workload blockwise_c1
step bench_c3       300 python bench.py --workload int8_c3
# 8-bit quantize with the cheaper (CPU-proven) search: the strict bit-exactness tests against the reference CUDA library
step q8fast_tests   300 env BNB_B200_Q8_FAST=1 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_zz_golden.py -q -k "8bit or None or quant"
step q8fast_bench   300 env BNB_B200_Q8_FAST=1 python bench.py --workload blockwise_c1 --no-cpu-baseline
# 4-bit GEMM variants: bit-equality with the default kernel, then timing
step gemm4_variants 400 python tools/probe_pair.py
step gemm4_persist  600 python tools/check_persistent.py
for f in gpurun_out/r2_gemm4_variants.log gpurun_out/r2_gemm4_persist.log gpurun_out/r2_bench.log gpurun_out/r2_bench_c1.log gpurun_out/r2_bench_c3.log gpurun_out/r2_q8fast_bench.log; do echo "---- $f"; tail -12 "$f"; done
