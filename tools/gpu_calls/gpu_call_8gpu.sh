#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 --workload sharded70b > gpurun_out/g${N}_sharded.log 2> gpurun_out/g${N}_sharded.err; echo "sharded rc=$?"; tail -c 400 gpurun_out/g${N}_sharded.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 50 --warmup 5 --no-secondary > gpurun_out/g${N}_bench.log 2> gpurun_out/g${N}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/g${N}_bench.err
