#!/usr/bin/env python
"""Golden vectors for the optimizer row, produced by running the REFERENCE itself.

Run here (the build container), never on the GPU box: imports the reference Python package from /root/reference
through the throw-away symlink package of make_golden.py (nothing of the reference is copied) and calls its own
CPU kernels of `bitsandbytes::optimizer_update_32bit` / `optimizer_update_8bit_blockwise` (reference
bitsandbytes/backends/cpu/ops.py:345-580: pure PyTorch, exact division / sqrt) on seeded inputs, three steps each.
Only the small .npz next to this script is committed; tests/test_optim_cpu.py pins oracle/optim_ref.py against it.

    python tests/golden/make_golden_optim.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402

HYPER = {  # lr, beta1, beta2, beta3, alpha, eps
    "adam": (1e-3, 0.9, 0.999, 0.0, 0.0, 1e-8),
    "momentum": (1e-2, 0.9, 0.0, 0.0, 0.0, 0.0),
    "rmsprop": (1e-2, 0.99, 0.0, 0.0, 0.0, 1e-8),
    "adagrad": (1e-2, 0.0, 0.0, 0.0, 0.0, 1e-10),
    "lion": (1e-4, 0.9, 0.99, 0.0, 0.0, 0.0),
    "ademamix": (1e-3, 0.9, 0.999, 0.9999, 5.0, 1e-8),
}


def main():
    bnb = import_reference(with_native=False)
    ops = torch.ops.bitsandbytes
    F = bnb.functional
    out = {}
    torch.manual_seed(0)
    n = 1000
    out["n"] = np.int64(n)
    code1 = F.create_dynamic_map(signed=True)
    code2 = F.create_dynamic_map(signed=False)
    out["code1"], out["code2"] = code1.numpy(), code2.numpy()
    for name, (lr, b1, b2, b3, alpha, eps) in HYPER.items():
        for wd in (0.0, 0.01):
            tag = f"{name}_wd{int(wd > 0)}"
            p = torch.randn(n) * 0.5
            grads = [torch.randn(n) * 0.1 for _ in range(3)]
            two = name in ("adam", "ademamix")
            out[f"{tag}_p0"] = p.numpy().copy()
            out[f"{tag}_g"] = torch.stack(grads).numpy().copy()
            # ---- 32-bit state
            p32 = p.clone()
            s1 = torch.zeros((2, n) if name == "ademamix" else (n,))
            s2 = torch.zeros(n) if two else None
            for step, g in enumerate(grads, 1):
                ops.optimizer_update_32bit(name, g.clone(), p32, s1, s2, None, 0.0, 0.0, b1, b2, b3, alpha, eps, wd, step, lr, 1.0,
                                           False)
                out[f"{tag}_32_p{step}"] = p32.numpy().copy()
                out[f"{tag}_32_s1_{step}"] = s1.numpy().copy()
                if two:
                    out[f"{tag}_32_s2_{step}"] = s2.numpy().copy()
            # ---- blockwise 8-bit state (one step from a random mid-training state; wd = 0 only: the CPU kernel
            # applies the decay before the update, the CUDA kernel after it)
            if wd == 0.0:
                nb = -(-n // 256)
                rows = 2 if name == "ademamix" else 1
                lo = 128 if name in ("rmsprop", "adagrad") else 0
                c1 = torch.randint(lo, 256, (rows, n) if rows == 2 else (n,), dtype=torch.uint8)
                c2 = torch.randint(0, 256, (n,), dtype=torch.uint8) if two else None
                a1 = torch.rand((rows, nb) if rows == 2 else (nb,)) * 0.05 + 1e-3
                a2 = torch.rand(nb) * 0.002 + 1e-5 if two else None
                out[f"{tag}_8_c1"], out[f"{tag}_8_a1"] = c1.numpy().copy(), a1.numpy().copy()
                if two:
                    out[f"{tag}_8_c2"], out[f"{tag}_8_a2"] = c2.numpy().copy(), a2.numpy().copy()
                p8 = p.clone()
                ops.optimizer_update_8bit_blockwise(name, grads[0].clone(), p8, c1, c2, b1, b2, b3, alpha, eps, 2, lr, code1,
                                                    code2 if two else None, a1, a2, wd, 1.0, False)
                out[f"{tag}_8_p"] = p8.numpy().copy()
                out[f"{tag}_8_c1_out"], out[f"{tag}_8_a1_out"] = c1.numpy().copy(), a1.numpy().copy()
                if two:
                    out[f"{tag}_8_c2_out"], out[f"{tag}_8_a2_out"] = c2.numpy().copy(), a2.numpy().copy()
    path = HERE / "reference_optim.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(out)} arrays)")


if __name__ == "__main__":
    main()
