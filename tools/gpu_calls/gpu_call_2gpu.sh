#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_full.py -q -k "two_gpu or two_devices or devices" > gpurun_out/g2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/g2_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/g2_bench.log 2> gpurun_out/g2_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/g2_bench.err
