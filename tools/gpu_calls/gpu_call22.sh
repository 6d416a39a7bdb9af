#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload optim_f4 > gpurun_out/c22_optim_bench.log 2> gpurun_out/c22_optim_bench.err; echo "rc=$?"; tail -c 400 gpurun_out/c22_optim_bench.err
