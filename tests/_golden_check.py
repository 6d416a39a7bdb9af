"""Shared checker: run a quantize / dequantize implementation over the reference-generated golden
vectors (tests/golden/reference_vectors.npz).  The GPU suite passes the C-ABI calls of the product
library; the CPU suite passes the oracle, which validates this checker itself.

quantize_fn(A_bits_or_f32: np.ndarray, dtype: str, blocksize, qt, code) -> (codes u8, absmax f32)
dequantize_fn(codes u8, absmax f32, blocksize, n, qt, code, out_dtype: str) -> np.ndarray of bit patterns
(uint16 for bf16/fp16, float32 viewed as uint32 for fp32)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "reference_vectors.npz"
Q8_NAMES = ["a", "b", "c", "d", "e", "f", "g", "h"]
Q4_NAMES = ["a", "b", "c", "d", "e", "f", "g", "h", "i"]
OUT_DTYPES = ["fp32", "bf16", "fp16"]


def load():
    return np.load(GOLDEN)


def _canon_zero(bits: np.ndarray) -> np.ndarray:
    """+0 and -0 compare equal (FP4 has two zeros: the CUDA kernels keep the sign, the CPU ones do not)."""
    b = bits.copy()
    if b.dtype == np.uint16:
        b[b == 0x8000] = 0
    else:
        b[b == 0x80000000] = 0
    return b


def _as_bits(x: np.ndarray) -> np.ndarray:
    return x.view(np.uint32) if x.dtype == np.float32 else x.view(np.uint16)


def check_8bit(z, name, quantize_fn, dequantize_fn):
    A = z[f"q8_{name}_A"]
    bs = int(z[f"q8_{name}_bs"])
    code = z["dynamic_map"]
    n = A.size
    codes, absmax = quantize_fn(A, "fp32", bs, None, code)
    np.testing.assert_array_equal(absmax, z[f"q8_{name}_absmax"], err_msg="absmax must be bit-exact")
    ref_codes = z[f"q8_{name}_codes"].reshape(-1)
    # codes: only inputs within a few ulp of a decision threshold may differ (GPU: MUFU.RCP normalisation)
    differing = int((codes.reshape(-1) != ref_codes).sum())
    assert differing <= max(2, n // 500), f"{differing} of {n} codes differ from the reference"
    for dt in OUT_DTYPES:
        got = dequantize_fn(ref_codes, z[f"q8_{name}_absmax"], bs, n, None, code, dt)
        np.testing.assert_array_equal(_as_bits(got).reshape(-1), _as_bits(z[f"q8_{name}_deq_{dt}_default"]).reshape(-1),
                                      err_msg=f"dequantize -> {dt} must be bit-exact")


def check_4bit(z, qt, name, quantize_fn, dequantize_fn):
    key = f"q4_{qt}_{name}"
    dt_in = str(z[f"{key}_dtype"])
    A = z[f"{key}_A"]
    bs = int(z[f"{key}_bs"])
    n = A.size
    packed, absmax = quantize_fn(A.reshape(-1), dt_in, bs, qt, None)
    np.testing.assert_array_equal(absmax, z[f"{key}_absmax"], err_msg="absmax must be bit-exact")
    ref = z[f"{key}_packed"].reshape(-1)
    assert packed.reshape(-1).shape == ref.shape

    def nibbles(p):
        hi, lo = p >> 4, p & 15
        if qt == "fp4":  # the two zeros are the same value
            hi = np.where(hi == 8, 0, hi)
            lo = np.where(lo == 8, 0, lo)
        return np.stack([hi, lo], axis=1).reshape(-1)[:n]

    differing = int((nibbles(packed.reshape(-1)) != nibbles(ref)).sum())
    assert differing <= max(1, n // 200), f"{differing} of {n} 4-bit codes differ from the reference"
    for dt in OUT_DTYPES:
        got = dequantize_fn(ref, z[f"{key}_absmax"], bs, n, qt, None, dt)
        want = z[f"{key}_deq_{dt}_default"].reshape(-1)
        np.testing.assert_array_equal(_canon_zero(_as_bits(got).reshape(-1)), _canon_zero(_as_bits(want)),
                                      err_msg=f"dequantize_4bit -> {dt} must be bit-exact")
