"""bench.py --workload sharded70b: BASELINE.json configs[3] -- matmul_4bit FP4 + double quant on
the Llama-3-70B FFN shape 8192 -> 28672, column-sharded across the ranks with an NCCL
all-gather of the partial outputs (strong scaling: the layer is fixed, ranks split it).

Per step every rank runs the fused kernel on its row shard of the globally quantised weight.
Three timings, max over ranks: the GEMM alone; GEMM + NCCL all_gather_into_tensor (the baseline
exchange); and the product path -- the GEMM epilogue stores its tile into every rank's
symmetric-memory output buffer over NVLink (parallel.fused_forward: no collective, one
symmetric-memory barrier per step).  `value` is the product path when it is available.
"""
from __future__ import annotations

import json
import time

N_FULL, K_FULL = 28672, 8192


def run_sharded70b(args, rank: int, world: int, local_rank: int) -> None:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = measure_sharded70b(dev, rank, world, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_sharded70b(dev, rank: int, world: int, steps: int, warmup: int):
    """All ranks call this (the process group is the caller's); returns the result line (meaningful on rank 0)."""
    import torch
    import torch.distributed as dist

    import bitsandbytes_b200.functional as F
    from bitsandbytes_b200.parallel import ColumnParallelLinear4bit, PeerGather, fused_forward, slice_quantized_weight

    class _A:
        pass

    args = _A()
    args.steps, args.warmup = steps, warmup
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
    ms_fused, fused_note = None, None
    if world > 1:
        try:
            peers = PeerGather(M, N_FULL, torch.bfloat16, dev)

            def gemm_fused(i):
                fused_forward(layer, xs[i % 3], peers)

            # parity of the exchange before timing it: fused result == NCCL-gathered result, bit for bit
            ref = layer(xs[0]).reshape(M, N_FULL).clone()
            got = fused_forward(layer, xs[0], peers)
            torch.cuda.synchronize()
            if not torch.equal(ref, got):
                raise RuntimeError("fused peer-store gather differs from the NCCL gather")
            ms_fused = timed(gemm_fused)
        except Exception as exc:  # noqa: BLE001  (no symmetric memory on this box: report and keep the NCCL number)
            fused_note = repr(exc)[:300]
    wall = time.perf_counter() - t0
    ms_value = ms_fused if ms_fused is not None else ms_all
    if True:
        line = {
            "metric": "fp4_dq_column_sharded_linear_tflops", "value": flops * args.steps / (ms_value * 1e-3) / 1e12,
            "unit": "TFLOPS", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "sharded70b", "N": N_FULL, "K": K_FULL, "M": M, "quant_type": "fp4",
                       "double_quant": True, "blocksize": 64, "parallelism": f"column-sharded x{world} + all-gather",
                       "rows_per_rank": shard.rows},
            "exchange": ("GEMM epilogue peer stores into symmetric memory (no collective)" if ms_fused is not None
                         else "NCCL all_gather_into_tensor"),
            "fused_peer_store_tflops": None if ms_fused is None else flops * args.steps / (ms_fused * 1e-3) / 1e12,
            "nccl_gather_tflops": flops * args.steps / (ms_all * 1e-3) / 1e12,
            "fused_unavailable": fused_note,
            "gemm_only_tflops": flops * args.steps / (ms_gemm * 1e-3) / 1e12,
            "gemm_only_ms_per_step": ms_gemm / args.steps,
            "gather_bytes_per_rank": 2 * M * shard.rows * (world - 1),
            "gpu_launches": args.steps, "wall_s": wall,
        }
    del layer, shard, qW, qs, xs, stage
    torch.cuda.empty_cache()
    return line
