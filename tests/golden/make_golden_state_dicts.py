"""Golden state dicts: what the REFERENCE serialises for a quantised Linear4bit / Linear8bitLt.

Run in this container (the reference is imported from /root/reference exactly as make_golden.py does,
"default" phase: no native library needed -- quantisation runs through its pure-torch kernels on
CPU):

    python tests/golden/make_golden_state_dicts.py      ->  tests/golden/reference_state_dicts.npz

Every tensor of every state dict is stored under "<case>::<state-dict key>", bit patterns for 16-bit
floats ("<...>::__dtype__<key>" records the torch dtype), plus the seeded float weights the layers
were built from, so tests/test_host_cpu.py can (a) load the reference's checkpoint into OUR modules
and get the same quantisation state, and (b) check that OUR modules serialise back to the same keys
and bytes: the drop-in boundary for `from_pretrained` checkpoints (reference nn/modules.py:213-484,
540-640, functional.py:447-600)."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402


def to_np(t: torch.Tensor):
    if t.dtype in (torch.bfloat16, torch.float16):
        return t.contiguous().view(torch.int16).numpy().view(np.uint16), str(t.dtype)
    return t.contiguous().numpy(), str(t.dtype)


def main():
    bnb = import_reference(with_native=False)
    out = {}
    K, N = 128, 96
    for case, qt, nested, storage in (("l4_nf4_plain", "nf4", False, torch.uint8),
                                      ("l4_nf4_nested", "nf4", True, torch.uint8),
                                      ("l4_fp4_plain", "fp4", False, torch.uint8),
                                      ("l4_nf4_bf16storage", "nf4", False, torch.bfloat16)):
        torch.manual_seed(hash(case) % 1000 if False else sum(map(ord, case)))
        lin = torch.nn.Linear(K, N, bias=True)
        m = bnb.nn.Linear4bit(K, N, bias=True, compute_dtype=torch.bfloat16, compress_statistics=nested, quant_type=qt,
                              quant_storage=storage)
        m.load_state_dict(lin.state_dict())
        m = m.to("cpu")
        if not m.weight.bnb_quantized:
            m.weight = m.weight._quantize(torch.device("cpu"))
        assert m.weight.bnb_quantized and m.weight.quant_state is not None
        out[f"{case}::__float_weight__"] = lin.weight.detach().numpy()
        out[f"{case}::__float_bias__"] = lin.bias.detach().numpy()
        for k, v in m.state_dict().items():
            arr, dt = to_np(v.detach())
            out[f"{case}::{k}"] = arr
            out[f"{case}::__dtype__{k}"] = np.array(dt)
        qs = m.weight.quant_state
        out[f"{case}::__meta__"] = np.array([qs.blocksize, int(qs.nested), K, N], np.int64)

    # LLM.int8(): int8 weight + SCB
    torch.manual_seed(77)
    lin = torch.nn.Linear(K, N, bias=True)
    m8 = bnb.nn.Linear8bitLt(K, N, bias=True, has_fp16_weights=False, threshold=6.0)
    m8.load_state_dict(lin.state_dict())
    try:
        m8.weight = m8.weight._quantize(torch.device("cpu")) if hasattr(m8.weight, "_quantize") else m8.weight
    except Exception as exc:  # noqa: BLE001
        print("int8 quantisation on CPU not available in this reference build:", exc)
    if m8.weight.dtype == torch.int8:
        out["l8::__float_weight__"] = lin.weight.detach().numpy()
        out["l8::__float_bias__"] = lin.bias.detach().numpy()
        for k, v in m8.state_dict().items():
            arr, dt = to_np(v.detach())
            out[f"l8::{k}"] = arr
            out[f"l8::__dtype__{k}"] = np.array(dt)
    out["reference_version"] = np.array(bnb.__version__)
    path = HERE / "reference_state_dicts.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(out)} arrays)")
    for k in sorted(out):
        if "::" in k and "__dtype__" not in k:
            print("  ", k, out[k].shape, out[k].dtype)


if __name__ == "__main__":
    main()
