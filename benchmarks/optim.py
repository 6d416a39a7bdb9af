"""bench.py --workload optim_f4: the optimizer updates of SURVEY.md section 8 row f-4 (HBM-bound element-wise
kernels): GB/s against the measured HBM copy peak, the reference CUDA library (oracle/_ref) on the same buffers
beside it.  Algorithmic bytes per element: g (read) + p (read + write) + the states (read + write) [+ absmax]."""
from __future__ import annotations

import ctypes as ct
import json


def _p(t):
    return None if t is None else ct.c_void_p(t.data_ptr())


def measure_optim(dev, ref, n: int = 64 * 1024 * 1024):
    import torch

    import bitsandbytes_b200.functional as F
    from benchmarks.common import hbm_peak, time_us
    from bitsandbytes_b200.cextension import lib

    peak, peak_src = hbm_peak()
    res = {}
    code1 = F.create_dynamic_map(signed=True).to(dev).contiguous()
    code2 = F.create_dynamic_map(signed=False).to(dev).contiguous()
    nb = -(-n // 256)
    f = ct.c_float
    for name, dtype, esz in (("adam", torch.bfloat16, 2), ("adam", torch.float32, 4), ("lion", torch.bfloat16, 2)):
        suf = {torch.bfloat16: "bf16", torch.float32: "fp32"}[dtype]
        two = name == "adam"
        p = (torch.randn(n, device=dev) * 0.1).to(dtype)
        g = (torch.randn(n, device=dev) * 0.01).to(dtype)
        # ---- blockwise 8-bit state
        c1 = torch.randint(0, 256, (n,), device=dev, dtype=torch.uint8)
        c2 = torch.randint(0, 256, (n,), device=dev, dtype=torch.uint8) if two else None
        a1 = torch.rand(nb, device=dev) * 0.05 + 1e-3
        a2 = torch.rand(nb, device=dev) * 0.002 + 1e-5 if two else None
        bytes8 = n * (3 * esz + (4 if two else 2)) + nb * (16 if two else 8)
        sym8 = f"c{name}_8bit_blockwise_grad_{suf}"

        def call8(L):
            fn = getattr(L, sym8)
            fn.restype = None
            fn.argtypes = [ct.c_void_p] * 4 + [f] * 5 + [ct.c_int32, f] + [ct.c_void_p] * 4 + [f, f, ct.c_bool, ct.c_int32]
            return lambda i: fn(_p(p), _p(g), _p(c1), _p(c2), 0.9, 0.999, 0.0, 0.0, 1e-8, 3 + i, 1e-4, _p(code1),
                                _p(code2) if two else None, _p(a1), _p(a2), 0.01, 1.0, False, n)

        # the reference-named symbols run on the legacy default stream: eager timing (not capturable)
        us, mode = time_us(call8(lib), 10, use_graph=False)
        entry = {"us": us, "gb_per_s": bytes8 / us / 1e3, "frac_of_hbm_peak": bytes8 / us / 1e3 / peak, "bytes": bytes8,
                 "elements": n, "timing": mode}
        if ref is not None:
            rus, _ = time_us(call8(ref), 10, use_graph=False)
            entry["ref_cuda_us"] = rus
        res[f"{name}_8bit_blockwise_{suf}_64Mi"] = entry
        del c1, c2
        # ---- 32-bit state
        s1 = torch.zeros(n, device=dev)
        s2 = torch.zeros(n, device=dev) if two else None
        bytes32 = n * (3 * esz + (16 if two else 8))
        sym32 = f"c{name}32bit_grad_{suf}"

        def call32(L):
            fn = getattr(L, sym32)
            fn.restype = None
            fn.argtypes = [ct.c_void_p] * 5 + [f] * 8 + [ct.c_int32, f, f, ct.c_bool, ct.c_int32]
            return lambda i: fn(_p(g), _p(p), _p(s1), _p(s2), None, 0.0, 0.0, 0.9, 0.999, 0.0, 0.0, 1e-8, 0.01, 3 + i, 1e-4, 1.0,
                                False, n)

        us, mode = time_us(call32(lib), 10, use_graph=False)
        entry = {"us": us, "gb_per_s": bytes32 / us / 1e3, "frac_of_hbm_peak": bytes32 / us / 1e3 / peak, "bytes": bytes32,
                 "elements": n, "timing": mode}
        if ref is not None:
            rus, _ = time_us(call32(ref), 10, use_graph=False)
            entry["ref_cuda_us"] = rus
        res[f"{name}_32bit_{suf}_64Mi"] = entry
        del s1, s2, p, g
        torch.cuda.empty_cache()
    return {"results": res, "hbm_peak_gbs": peak, "peak_source": peak_src}


def run_optim_f4(args, rank, world, local_rank):
    import torch

    from benchmarks.paths import load_ref_cuda

    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    r = measure_optim(dev, load_ref_cuda())
    head = r["results"]["adam_8bit_blockwise_bf16_64Mi"]
    line = {"metric": "adam8bit_blockwise_update_gb_per_s", "value": head["gb_per_s"], "unit": "GB/s", "n_gpus": 1,
            "steps": 10, "warmup": 3, "ms_per_step": head["us"] * 1e-3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 parameters, 8-bit blockwise state", "data": "synthetic (randn)",
            "config": {"workload": "optim_f4", "elements": head["elements"], "optimizer": "adam", "l2": "64 Mi elements per tensor (> L2)"},
            "roofline": {"bound": "hbm", "achieved": head["gb_per_s"], "peak": r["hbm_peak_gbs"], "unit": "GB/s",
                         "frac": head["frac_of_hbm_peak"], "traffic": None, "peak_source": r["peak_source"]},
            "results": r["results"], "gpu_launches": 10}
    print(json.dumps(line), flush=True)
