#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/c25_tests.log 2>&1; tail -4 gpurun_out/c25_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary > gpurun_out/c25_bench.log 2> gpurun_out/c25_bench.err; echo "bench rc=$?"; python -c "
import json
for l in open('gpurun_out/c25_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']), 'launches', d['gpu_launches'])
"
