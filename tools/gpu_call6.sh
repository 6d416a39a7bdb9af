#!/bin/bash
mkdir -p gpurun_out
step() { name=$1; shift; echo "== $name"; timeout "$1" "${@:2}" > "gpurun_out/c6_$name.log" 2>&1; echo "   rc=$? ($(tail -1 gpurun_out/c6_$name.log | cut -c1-160))"; }
step lite   300 python tools/probe_gemm4_pair.py lite
step bench_c1 300 python bench.py --workload blockwise_c1 --no-cpu-baseline
step q8_tests 300 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_zz_golden.py -q
cat gpurun_out/c6_lite.log | grep lite
tail -3 gpurun_out/c6_q8_tests.log
python - <<'PY'
import json
for line in open("gpurun_out/c6_bench_c1.log"):
    if line.startswith("{"):
        d=json.loads(line)
        print({k:(round(v["us"],1),round(v["frac_of_hbm_peak"],2)) for k,v in d["results"].items()})
PY
