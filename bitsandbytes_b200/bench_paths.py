"""bench.py workloads for the two BASELINE.json configs the default line does not cover:

``--workload blockwise_c1``  configs[0] on the GPU: quantize_blockwise / dequantize_blockwise of a
    4 Mi-element fp32 tensor (blocksize 4096 = the API default, and 256), GB/s of ALGORITHMIC
    bytes (``4 n + n + 4 n / bs``) against the measured HBM copy peak.
``--workload int8_c3``       configs[2]: ``Linear8bitLt`` (LLM.int8(), threshold 6.0) forward at the
    Llama-3-8B FFN shape 4096 -> 11008, 4096 tokens, fp16, with the reference benchmark's five
    outlier columns (reference benchmarking/matmul_benchmark.py:47-48), through ``module.forward``.

Both run on rank 0 of one GPU (they do not shard) and time with CUDA events over rotating buffer sets
larger than L2.  ``cpu_baseline`` is a callback supplied by bench.py (the oracle port on a bounded
sample): this package never imports ``oracle/``.
"""
from __future__ import annotations

import json
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
L2_BYTES = 126 * 1024 * 1024


def _hbm_peak():
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:  # noqa: BLE001
        return 7700.0, "B200_PROFILING.md fallback (nominal HBM3e)"


def _time_us(fn, steps, warmup):
    """Mean microseconds per call of fn(i) over `steps` back-to-back launches (CUDA events)."""
    import torch

    for i in range(max(warmup, 3)):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / steps


def run_blockwise_c1(args, rank: int, world: int, local_rank: int, cpu_baseline=None) -> None:
    import torch

    from . import functional as F

    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = 4 * 1024 * 1024
    torch.manual_seed(0)
    code = F.create_dynamic_map().to(dev)
    sets = max(3, int(2 * L2_BYTES // (n * 9)) + 1)  # fp32 in + uint8 codes + fp32 out per set
    As = [torch.randn(n, device=dev) for _ in range(sets)]
    peak, peak_src = _hbm_peak()
    results = {}
    launches = 0
    for bs in (4096, 256):
        algo_bytes = 4 * n + n + 4 * n // bs
        qs = [F.quantize_blockwise(a, code=code, blocksize=bs) for a in As]
        outs = [torch.empty(n, device=dev) for _ in range(sets)]

        def quant(i, bs=bs):
            F.quantize_blockwise(As[i % sets], code=code, blocksize=bs)

        def dequant(i, bs=bs, qs=qs, outs=outs):
            q, st = qs[i % sets]
            F.dequantize_blockwise(q, st, out=outs[i % sets])

        for name, fn in (("quantize", quant), ("dequantize", dequant)):
            us = _time_us(fn, args.steps, args.warmup)
            launches += args.steps
            gbs = algo_bytes / us / 1e3
            results[f"{name}_bs{bs}"] = {"us": us, "gb_per_s": gbs, "frac_of_hbm_peak": gbs / peak,
                                         "algorithmic_bytes": algo_bytes}

    # CPU baseline (supplied by bench.py: the package itself never touches oracle/)
    cpu = cpu_baseline(n, code.cpu().numpy()) if cpu_baseline is not None else None

    head = results["dequantize_bs4096"]
    line = {
        "metric": "blockwise_dequantize_fp32_gb_per_s", "value": head["gb_per_s"], "unit": "GB/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["us"] * 1e-3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (randn, seed 0)",
        "config": {"workload": "blockwise_c1", "elements": n, "blocksizes": [4096, 256], "code": "dynamic map (8-bit)",
                   "l2": f"rotating {sets} buffer sets ({sets * n * 9 / 2**20:.0f} MiB > 126 MiB L2)"},
        "results": results,
        "roofline": {"bound": "hbm", "achieved": head["gb_per_s"], "peak": peak, "unit": "GB/s",
                     "frac": head["gb_per_s"] / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "dequantize_blockwise_kernel<float, 8-bit>"},
        "cpu_baseline": cpu, "gpu_launches": launches,
        "e2e": None,
    }
    print(json.dumps(line), flush=True)


def run_int8_c3(args, rank: int, world: int, local_rank: int, cpu_baseline=None) -> None:
    import torch

    import bitsandbytes_b200 as bnb

    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    K, N, M = 4096, 11008, 4096
    torch.manual_seed(0)
    lin = torch.nn.Linear(K, N, bias=False)
    layer = bnb.nn.Linear8bitLt(K, N, bias=False, has_fp16_weights=False, threshold=6.0)
    layer.load_state_dict(lin.state_dict())
    layer = layer.to(dev).eval()  # quantises the weight to int8 + SCB on the move to the device
    outlier_cols = torch.randint(0, K, (5,), generator=torch.Generator().manual_seed(1)).tolist()
    sets = 3  # activations 32 MiB + outputs 86 MiB per set: 3 sets exceed the 126 MiB L2
    xs = []
    for i in range(sets):
        x = torch.randn(M, K, device=dev, dtype=torch.float16)
        x[:, outlier_cols] = 8.0
        xs.append(x)
    steps = max(10, min(args.steps, 200))

    with torch.no_grad():
        def fwd(i):
            return layer(xs[i % sets])

        us = _time_us(fwd, steps, args.warmup)
    ops = 2.0 * M * N * K
    peak_tops = 4500.0  # nominal dense int8 (B200_PROFILING.md); no measured int8 figure in MEASURED_PEAKS.json

    cpu = None
    if cpu_baseline is not None:
        CB = layer.state.CB if layer.state.CB is not None else layer.weight.CB
        SCB = layer.state.SCB if layer.state.SCB is not None else layer.weight.SCB
        cpu = cpu_baseline(xs[0].cpu(), CB.cpu().numpy(), SCB.float().cpu().numpy(), M, N, K)

    line = {
        "metric": "linear8bitlt_forward_tops", "value": ops / us / 1e6, "unit": "TOPS", "n_gpus": 1, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": us * 1e-3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8 (fp16 activations, fp16 outlier columns)", "data": "synthetic (randn, seed 0)",
        "config": {"workload": "int8_c3", "N": N, "K": K, "M": M, "threshold": 6.0, "outlier_columns": len(set(outlier_cols)),
                   "api": "bitsandbytes_b200.nn.Linear8bitLt.forward",
                   "l2": f"rotating {sets} activation / output sets"},
        "tokens_per_s": M / (us * 1e-6),
        "roofline": {"bound": "tensor", "achieved": ops / us / 1e6, "peak": peak_tops, "unit": "TOP/s",
                     "frac": ops / us / 1e6 / peak_tops, "traffic": None,
                     "peak_source": "nominal dense int8 (B200_PROFILING.md)",
                     "kernel": "int8_gemm_tc_kernel (fused dequant epilogue) inside the module forward: the timed "
                               "step also holds the row quantisation, the outlier gather and the fp16 addmm"},
        "cpu_baseline": cpu, "gpu_launches": None, "e2e": None,
    }
    print(json.dumps(line), flush=True)
