#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c23_tests.log 2>&1; tail -3 gpurun_out/c23_tests.log
timeout 300 python bench.py --workload optim_f4 > gpurun_out/c23_optim_bench.log 2> gpurun_out/c23_optim_bench.err; echo "optim rc=$?"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/c23_bench.log 2> gpurun_out/c23_bench.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
