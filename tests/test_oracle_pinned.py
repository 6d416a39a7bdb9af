"""Pin the CPU oracle (oracle/oracle_c.c) against the reference's own outputs.

Golden vectors come from tests/golden/make_golden.py (the reference Python package run in
the build container, both its native C++ CPU backend and its pure-PyTorch "default"
kernels).  When the reference CPU library built from the reference sources is present
(oracle/_ref/libbitsandbytes_cpu*.so) it is additionally called directly through ctypes.

Bar: bit-exact for absmax, dequantize (fp32/bf16/fp16), the int8 GEMM, int8 dequant; for
quantization codes bit-exact except inputs that sit within 2 ulp of a decision threshold
(the reference's own implementations disagree there: `x * (1/absmax)` vs `x / absmax`).
"""
import ctypes as ct

import numpy as np
import pytest

import oracle


def _near_threshold_8bit(x_norm, code):
    bounds = (code[:-1].astype(np.float64) + code[1:].astype(np.float64)) / 2
    d = np.min(np.abs(x_norm.astype(np.float64)[:, None] - bounds[None, :]), axis=1)
    return d <= 4 * np.spacing(np.abs(x_norm).astype(np.float32)).astype(np.float64) + 1e-12


def _assert_bits_equal(got, want, dtype, fp4_zero=False):
    """Bit-exact comparison.  fp4_zero: FP4 code 8 is -0.0 in the CUDA kernels (lut[0] * -1,
    reference kernels.cu:59-62) but +0.0 in the CPU backend / torch table; fold the sign of zero."""
    if dtype == "fp32":
        got, want = got.view(np.uint32).copy(), want.view(np.uint32).copy()
        neg0 = np.uint32(0x80000000)
    else:
        got, want = got.copy(), want.copy()
        neg0 = np.uint16(0x8000)
    if fp4_zero:
        got[got == neg0] = 0
        want[want == neg0] = 0
    np.testing.assert_array_equal(got, want)


# cases of tests/golden/make_golden.py: a-c the original ones, d-i edge sizes (1 element, one short of a block,
# ragged tails, blocksize 32 ... 4096)
Q8_NAMES = ["a", "b", "c", "d", "e", "f", "g", "h"]
Q4_NAMES = ["a", "b", "c", "d", "e", "f", "g", "h", "i"]


def test_codebooks(golden):
    np.testing.assert_array_equal(oracle.lut4("nf4"), golden["nf4_code"])
    np.testing.assert_array_equal(oracle.lut4("fp4"), golden["fp4_code"])
    # -0.0 for FP4 code 8 (sign bit set on zero), as the reference builds it
    assert np.signbit(oracle.lut4("fp4")[8])


@pytest.mark.parametrize("name", Q8_NAMES)
def test_quantize_8bit_vs_reference_default(golden, name):
    A = golden[f"q8_{name}_A"]
    bs = int(golden[f"q8_{name}_bs"])
    code = golden["dynamic_map"]
    q, absmax = oracle.quantize_blockwise(A, bs, None, code)
    np.testing.assert_array_equal(absmax, golden[f"q8_{name}_absmax"])
    np.testing.assert_array_equal(absmax, golden[f"q8_{name}_absmax_cpulib"])
    ref = golden[f"q8_{name}_codes"].reshape(-1)
    bad = np.nonzero(q != ref)[0]
    if bad.size:
        x = A[bad] * (np.float32(1.0) / absmax[bad // bs])
        assert np.all(np.abs(q[bad].astype(int) - ref[bad].astype(int)) == 1)
        assert np.all(_near_threshold_8bit(x, code)), "mismatch away from a decision threshold"
    assert bad.size <= max(2, A.size // 5000)


@pytest.mark.parametrize("name", Q8_NAMES)
@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("which", ["default", "native"])
def test_dequantize_8bit_bit_exact(golden, name, dtype, which):
    bs = int(golden[f"q8_{name}_bs"])
    code = golden["dynamic_map"]
    codes = golden[f"q8_{name}_codes"] if which == "default" else golden[f"q8_{name}_codes_cpulib"]
    absmax = golden[f"q8_{name}_absmax"]
    got = oracle.dequantize_blockwise(codes, absmax, bs, codes.size, None, code, dtype)
    want = golden[f"q8_{name}_deq_{dtype}_{which}"].reshape(-1)
    if dtype == "fp32":
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    else:
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("qt", ["nf4", "fp4"])
@pytest.mark.parametrize("name", Q4_NAMES)
def test_quantize_4bit_vs_reference_default(golden, qt, name):
    key = f"q4_{qt}_{name}"
    dt = str(golden[f"{key}_dtype"])
    A = oracle.widen(golden[f"{key}_A"].reshape(-1), dt)
    bs = int(golden[f"{key}_bs"])
    packed, absmax = oracle.quantize_blockwise(A, bs, qt)
    np.testing.assert_array_equal(absmax, golden[f"{key}_absmax"])
    ref = golden[f"{key}_packed"]
    assert packed.shape == ref.shape
    if qt == "fp4":
        # FP4 has two zeros.  The CUDA tree keeps the sign of a value that rounds to zero
        # (code 8 = -0.0, reference kernels.cu:84), the torch "default" kernel emits code 0.
        # Both dequantize to zero; canonicalise before comparing.
        def canon(p):
            hi, lo = p >> 4, p & 15
            hi = np.where(hi == 8, 0, hi)
            lo = np.where(lo == 8, 0, lo)
            return (hi << 4) | lo

        packed, ref = canon(packed), canon(ref)
    bad = np.nonzero(packed != ref)[0]
    dist = oracle.lib().oracle_nf4_threshold_distance if qt == "nf4" else oracle.lib().oracle_fp4_threshold_distance
    for b in bad:
        for e in (2 * b, 2 * b + 1):
            if e >= A.size:
                continue
            x = np.float32(A[e]) * (np.float32(1.0) / absmax[e // bs])
            got = (packed[b] >> 4) if e % 2 == 0 else (packed[b] & 15)
            want = (ref[b] >> 4) if e % 2 == 0 else (ref[b] & 15)
            if got != want:
                # FP4: the CUDA tree's literals are 6-7 digit decimals (0.583333f vs the exact
                # midpoint 0.58333334 the torch kernel uses, reference kernels.cu:85-105)
                tol = 1e-6 if qt == "fp4" else 4 * np.spacing(np.float32(abs(x)))
                assert dist(float(x)) <= tol, (e, x, got, want)
    assert bad.size <= max(1, A.size // (500 if qt == "fp4" else 2000))


@pytest.mark.parametrize("qt", ["nf4", "fp4"])
@pytest.mark.parametrize("name", Q4_NAMES)
@pytest.mark.parametrize("dtype", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("which", ["default", "native"])
def test_dequantize_4bit_bit_exact(golden, qt, name, dtype, which):
    key = f"q4_{qt}_{name}"
    bs = int(golden[f"{key}_bs"])
    want = golden[f"{key}_deq_{dtype}_{which}"].reshape(-1)
    got = oracle.dequantize_blockwise(golden[f"{key}_packed"], golden[f"{key}_absmax"], bs, want.size, qt, None, dtype)
    _assert_bits_equal(got, want, dtype, fp4_zero=(qt == "fp4"))


@pytest.mark.parametrize("name", ["plain", "nested", "fp16"])
def test_gemm_4bit_vs_reference_public_api(golden, name):
    key = f"gemm4_{name}"
    M, N, K = (int(v) for v in golden[f"{key}_shape"])
    qt = str(golden[f"{key}_qt"])
    dt = str(golden[f"{key}_dtype"])
    x = oracle.widen(golden[f"{key}_x"].reshape(-1), dt)
    kw = {}
    if f"{key}_absmax8" in golden:
        kw = dict(absmax_8bit=golden[f"{key}_absmax8"], absmax_code=golden[f"{key}_code2"],
                  absmax_offset=float(golden[f"{key}_offset"][0]))
        absmax = golden[f"{key}_absmax2"]
        # F.dequantize_4bit's weights: the nested scale must reproduce bit-exactly
        scale = oracle.nested_absmax(absmax, kw["absmax_8bit"], kw["absmax_code"], kw["absmax_offset"])
        wdq = oracle.dequantize_blockwise(golden[f"{key}_packed"], scale, 64, N * K, qt, None, dt)
    else:
        absmax = golden[f"{key}_absmax"]
        wdq = oracle.dequantize_blockwise(golden[f"{key}_packed"], absmax, 64, N * K, qt, None, dt)
    _assert_bits_equal(wdq, golden[f"{key}_Wdq"], dt, fp4_zero=(qt == "fp4"))
    bias = oracle.widen(golden[f"{key}_bias"], dt) if f"{key}_bias" in golden else None
    y64 = oracle.gemm_4bit(x, golden[f"{key}_packed"], absmax, M, N, K, 64, qt, dt, bias, **kw)
    y_ref = oracle.widen(golden[f"{key}_y"], dt).reshape(M, N).astype(np.float64)
    # the reference CPU path = F.linear(x, Wdq) in T: |diff| within one T-ulp of the exact sum
    eps = 2.0**-8 if dt == "bf16" else 2.0**-11
    assert np.all(np.abs(y64 - y_ref) <= eps * np.abs(y64) + 2e-3)


@pytest.mark.parametrize("tag,thr", [("t0", 0.0), ("t6", 6.0)])
def test_int8_vector_quant(golden, tag, thr):
    A = golden[f"i8vq_{tag}_A"]
    q, stats = oracle.int8_vector_quant(A, thr)
    np.testing.assert_array_equal(stats, golden[f"i8vq_{tag}_stats"])
    ref = golden[f"i8vq_{tag}_q"].copy()
    cols = golden[f"i8vq_{tag}_cols"]
    # the reference zeroes whole outlier columns in Python after the kernel (cuda/ops.py:251-252)
    q2 = q.copy()
    if cols.size:
        q2[:, cols] = 0
        Af = oracle.widen(A, "fp16")
        np.testing.assert_array_equal(cols, np.nonzero((np.abs(Af) >= thr).any(axis=0))[0])
    assert np.max(np.abs(q2.astype(int) - ref.astype(int))) <= 1
    assert np.mean(q2 != ref) < 0.01


def test_int8_gemm_and_dequant(golden):
    C = oracle.int8_gemm(golden["i8mm_A"], golden["i8mm_B"])
    np.testing.assert_array_equal(C, golden["i8mm_C"])
    for tag, bias in (("nobias", None), ("bias", golden["i8mm_bias"])):
        got = oracle.widen(oracle.int8_mm_dequant(C, golden["i8mm_rs"], golden["i8mm_cs"], bias), "fp16")
        want = oracle.widen(golden[f"i8mm_deq_{tag}"], "fp16")
        # torch default impl multiplies by 6.200124e-05 in a different order (default/ops.py:57):
        # equal up to one fp16 ulp
        assert np.all(np.abs(got - want) <= np.spacing(np.abs(want).astype(np.float16)).astype(np.float32) + 1e-7)


def test_reference_cpu_library_direct():
    """Call the reference C++ CPU backend (built from the reference sources) through ctypes."""
    path = oracle.ref_cpu_library_path()
    if path is None:
        pytest.skip("oracle/_ref reference CPU library not built")
    ref = ct.CDLL(str(path))
    rng = np.random.default_rng(0)
    n, bs = 64 * 333, 64
    packed = rng.integers(0, 256, n // 2, dtype=np.uint8)
    absmax = (rng.random(n // bs, dtype=np.float32) * 3 + 0.01).astype(np.float32)
    for qt, sym in (("nf4", "nf4"), ("fp4", "fp4")):
        for dtype, suffix, npdt in (("fp32", "fp32", np.float32), ("bf16", "bf16", np.uint16), ("fp16", "fp16", np.uint16)):
            out = np.zeros(n, npdt)
            fn = getattr(ref, f"cdequantize_blockwise_cpu_{sym}_{suffix}")
            fn.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_longlong, ct.c_longlong, ct.c_longlong]
            fn(packed.ctypes.data, absmax.ctypes.data, out.ctypes.data, bs, n // 64, 64)
            want = oracle.dequantize_blockwise(packed, absmax, bs, n, qt, None, dtype)
            _assert_bits_equal(out, want, dtype, fp4_zero=(qt == "fp4"))


# The same checker the GPU suite runs over the golden vectors (tests/_golden_check.py), here with the
# oracle as the implementation: validates the checker and pins the oracle once more.
def _oracle_quantize(A, dtype, bs, qt, code):
    a = A if dtype == "fp32" else oracle.widen(A.reshape(-1), dtype)
    return oracle.quantize_blockwise(np.ascontiguousarray(a, dtype=np.float32), bs, qt, code)


def _oracle_dequantize(codes, absmax, bs, n, qt, code, out_dtype):
    return oracle.dequantize_blockwise(codes, absmax, bs, n, qt, code, out_dtype)


@pytest.mark.parametrize("name", Q8_NAMES)
def test_golden_checker_8bit_with_the_oracle(name):
    from tests import _golden_check as gc

    gc.check_8bit(gc.load(), name, _oracle_quantize, _oracle_dequantize)


@pytest.mark.parametrize("qt", ["nf4", "fp4"])
@pytest.mark.parametrize("name", Q4_NAMES)
def test_golden_checker_4bit_with_the_oracle(qt, name):
    from tests import _golden_check as gc

    gc.check_4bit(gc.load(), qt, name, _oracle_quantize, _oracle_dequantize)


def _oracle_gemm4(x_bits, dt, packed, absmax, a8, code2, offset, bias_bits, M, N, K, bs, qt):
    kw = {}
    if a8 is not None:
        kw = dict(absmax_8bit=a8, absmax_code=code2, absmax_offset=offset)
    bias = oracle.widen(bias_bits, dt) if bias_bits is not None else None
    y64 = oracle.gemm_4bit(oracle.widen(x_bits, dt), packed, absmax, M, N, K, bs, qt, dt, bias, **kw)
    return oracle.widen(oracle.round_to(np.asarray(y64, dtype=np.float32), dt), dt).astype(np.float64)


@pytest.mark.parametrize("name", ["plain", "nested", "fp16"])
def test_golden_checker_gemm4_with_the_oracle(name):
    from tests import _golden_check as gc

    gc.check_gemm4(gc.load(), name, _oracle_gemm4)


def test_golden_checker_int8_with_the_oracle():
    from tests import _golden_check as gc

    gc.check_int8_gemm(gc.load(), oracle.int8_gemm, oracle.int8_mm_dequant)
