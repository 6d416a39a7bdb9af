#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/c4_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/c4_$name.log | cut -c1-160))"; }
step pair_trace  300 python tools/probe_gemm4_pair.py trace
step pair_time   500 python tools/probe_gemm4_pair.py time 4096x4096x4096 4096x11008x4096 4096x4096x11008 1024x4096x4096 8192x8192x8192
step tests       1500 python -m pytest tests -q -m gpu
step decode      300 python tools/probe_decode.py
step bench_c1    300 python bench.py --workload blockwise_c1 --no-cpu-baseline
step bench_c1f   300 env BNB_B200_Q8_FAST=1 python bench.py --workload blockwise_c1 --no-cpu-baseline
step q8fast_tests 300 env BNB_B200_Q8_FAST=1 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_zz_golden.py -q -k "8bit or None or quant"
for f in gpurun_out/c4_pair_time.log gpurun_out/c4_tests.log gpurun_out/c4_decode.log gpurun_out/c4_q8fast_tests.log; do echo "---- $f"; tail -14 "$f" | cut -c1-900; done
grep -E "=====|MMA stage period|MMA thread|issue deltas|epilogue begin|decode:" gpurun_out/c4_pair_trace.log
python - <<'PY'
import json
for f in ("gpurun_out/c4_bench_c1.log","gpurun_out/c4_bench_c1f.log"):
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line)
            print(f, {k:(round(v["us"],1),round(v["frac_of_hbm_peak"],2)) for k,v in d["results"].items()})
PY
