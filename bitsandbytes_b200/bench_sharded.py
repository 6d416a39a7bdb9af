"""bench.py --workload sharded70b: BASELINE.json configs[3] -- matmul_4bit FP4 + double quant on
the Llama-3-70B FFN shape 8192 -> 28672, column-sharded across the ranks with an NCCL
all-gather of the partial outputs (strong scaling: the layer is fixed, ranks split it).

Per step every rank runs the fused kernel on its row shard of the globally quantised weight
(written straight into its slot of the gather buffer) and one all_gather_into_tensor exchanges
the slices.  Reported: whole-layer TFLOPS with and without the gather, max over ranks.
"""
from __future__ import annotations

import json
import time

N_FULL, K_FULL = 28672, 8192


def run_sharded70b(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist

    from . import functional as F
    from .parallel import ColumnParallelLinear4bit, slice_quantized_weight

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    M = 4096
    torch.manual_seed(0)  # identical weight on every rank: quantise once "globally", then slice
    W = (torch.randn(N_FULL, K_FULL, device=dev) / K_FULL**0.5).to(torch.bfloat16)
    qW, qs = F.quantize_4bit(W, blocksize=64, quant_type="fp4", compress_statistics=True)
    del W
    shard = slice_quantized_weight(qW, qs, world, rank)
    layer = ColumnParallelLinear4bit(shard, N_FULL)
    xs = [torch.randn(M, K_FULL, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    stage = torch.empty((world, M, shard.rows), device=dev, dtype=torch.bfloat16)
    flops = 2.0 * M * N_FULL * K_FULL

    def gemm_only(i):
        layer.local_forward(xs[i % 3], stage[rank], shard.rows)

    def gemm_gather(i):
        layer.local_forward(xs[i % 3], stage[rank], shard.rows)
        if world > 1:
            dist.all_gather_into_tensor(stage.view(-1), stage[rank].reshape(-1))

    def timed(fn):
        for i in range(max(args.warmup, 3)):
            fn(i)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            fn(i)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t0 = time.perf_counter()
    ms_gemm = timed(gemm_only)
    ms_all = timed(gemm_gather)
    wall = time.perf_counter() - t0
    if rank == 0:
        line = {
            "metric": "fp4_dq_column_sharded_linear_tflops", "value": flops * args.steps / (ms_all * 1e-3) / 1e12,
            "unit": "TFLOPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "sharded70b", "N": N_FULL, "K": K_FULL, "M": M, "quant_type": "fp4",
                       "double_quant": True, "blocksize": 64, "parallelism": f"column-sharded x{world} + all-gather",
                       "rows_per_rank": shard.rows},
            "gemm_only_tflops": flops * args.steps / (ms_gemm * 1e-3) / 1e12,
            "gemm_only_ms_per_step": ms_gemm / args.steps,
            "gather_bytes_per_rank": 2 * M * shard.rows * (world - 1),
            "gpu_launches": args.steps, "wall_s": wall,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
