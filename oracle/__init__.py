"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's hot-path algorithms (``oracle_c.c``, plain C, each
function citing the reference file:line it follows) plus loaders for the two reference
libraries built from the reference's own sources by ``oracle/Makefile`` into
``oracle/_ref/`` (CPU backend and CUDA backend; see the Makefile header).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs -- always as the checker or the reported
baseline, never on the product path.  ``bitsandbytes_b200`` never imports it.

Parity pinning status: PINNED.  ``tests/test_oracle_pinned.py`` checks this restatement
against (a) golden vectors produced by importing the reference Python package
(``tests/golden/make_golden.py``) and (b) the reference CPU library compiled from the
reference sources; the GPU tests additionally run the reference CUDA library
(``_ref/libbitsandbytes_cuda_ref.so``) on the same device buffers.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).parent
REF_DIR = HERE / "_ref"

GENERAL8BIT, FP4, NF4 = 0, 1, 2
QUANT_TYPE = {"fp4": FP4, "nf4": NF4, None: GENERAL8BIT, "general8bit": GENERAL8BIT}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_L = ct.c_long


def build(ref: bool = False, quiet: bool = True) -> None:
    """Compile the C restatement (and, if asked and /root/reference exists, the reference libs)."""
    targets = ["port"]
    if ref and Path(os.environ.get("BNB_REFERENCE_DIR", "/root/reference")).exists():
        targets.append("ref")
    subprocess.run(
        ["make", "-C", str(HERE)] + targets,
        check=True,
        stdout=subprocess.DEVNULL if quiet else None,
        stderr=subprocess.STDOUT if quiet else None,
    )


_lib = None


def lib() -> ct.CDLL:
    global _lib
    if _lib is None:
        path = REF_DIR / "liboracle_c.so"
        if not path.exists():
            build()
        L = ct.CDLL(str(path))
        L.oracle_get_4bit_lut.argtypes = [_f32p, ct.c_int]
        L.oracle_quantize_blockwise.argtypes = [ct.c_void_p, _f32p, _f32p, _u8p, _L, _L, ct.c_int]
        L.oracle_dequantize_blockwise.argtypes = [ct.c_void_p, _u8p, _f32p, _f32p, _L, _L, ct.c_int]
        L.oracle_round_bf16.argtypes = [_f32p, _u16p, _L]
        L.oracle_round_fp16.argtypes = [_f32p, _u16p, _L]
        L.oracle_widen_bf16.argtypes = [_u16p, _f32p, _L]
        L.oracle_widen_fp16.argtypes = [_u16p, _f32p, _L]
        L.oracle_gemm_4bit.argtypes = [_f32p, _u8p, _f32p, ct.c_void_p, ct.c_void_p, ct.c_void_p, _f64p,
                                       ct.c_void_p, _L, _L, _L, _L, ct.c_int, ct.c_int]
        L.oracle_nested_absmax.argtypes = [_f32p, _u8p, _f32p, ct.c_float, _f32p, _L]
        L.oracle_int8_vector_quant.argtypes = [_u16p, _i8p, _f32p, ct.c_float, _L, _L]
        L.oracle_int8_gemm.argtypes = [_i8p, _i8p, _i32p, _L, _L, _L]
        L.oracle_int8_mm_dequant.argtypes = [_i32p, _f32p, _f32p, _u16p, ct.c_void_p, _L, _L]
        L.oracle_nf4_threshold_distance.argtypes = [ct.c_float]
        L.oracle_nf4_threshold_distance.restype = ct.c_float
        L.oracle_fp4_threshold_distance.argtypes = [ct.c_float]
        L.oracle_fp4_threshold_distance.restype = ct.c_float
        L.oracle_num_threads.restype = ct.c_int
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ct.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------------------
# dtype helpers: 16-bit floats travel as uint16 bit patterns
# ----------------------------------------------------------------------------------------
def widen(bits: np.ndarray, dtype: str) -> np.ndarray:
    """bf16/fp16 bit patterns (uint16) -> exact fp32."""
    if dtype == "fp32":
        return _c(bits, np.float32)
    bits = _c(bits, np.uint16)
    out = np.empty(bits.shape, np.float32)
    (lib().oracle_widen_bf16 if dtype == "bf16" else lib().oracle_widen_fp16)(bits.reshape(-1), out.reshape(-1), bits.size)
    return out


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """fp32 -> T with round-to-nearest-even; returns uint16 bit patterns (or fp32 unchanged)."""
    x = _c(x, np.float32)
    if dtype == "fp32":
        return x
    out = np.empty(x.shape, np.uint16)
    (lib().oracle_round_bf16 if dtype == "bf16" else lib().oracle_round_fp16)(x.reshape(-1), out.reshape(-1), x.size)
    return out


def lut4(quant_type: str) -> np.ndarray:
    out = np.empty(16, np.float32)
    lib().oracle_get_4bit_lut(out, QUANT_TYPE[quant_type])
    return out


# ----------------------------------------------------------------------------------------
# blockwise quantize / dequantize
# ----------------------------------------------------------------------------------------
def quantize_blockwise(A_f32: np.ndarray, blocksize: int, quant_type=None, code: np.ndarray | None = None):
    """Returns (codes uint8, absmax fp32).  4-bit: two codes per byte, element 2b in the high nibble."""
    qt = QUANT_TYPE[quant_type]
    A = _c(A_f32, np.float32).reshape(-1)
    n = A.size
    nblocks = -(n // -blocksize)
    absmax = np.empty(nblocks, np.float32)
    out = np.empty(n if qt == 0 else (n + 1) // 2, np.uint8)
    code_c = _c(code, np.float32) if code is not None else None
    lib().oracle_quantize_blockwise(_ptr(code_c), A, absmax, out, blocksize, n, qt)
    return out, absmax


def dequantize_blockwise(codes: np.ndarray, absmax: np.ndarray, blocksize: int, n: int, quant_type=None,
                         code: np.ndarray | None = None, dtype: str = "fp32"):
    """Returns fp32 (dtype fp32) or uint16 bit patterns of T (bf16/fp16)."""
    qt = QUANT_TYPE[quant_type]
    out = np.empty(n, np.float32)
    code_c = _c(code, np.float32) if code is not None else None
    lib().oracle_dequantize_blockwise(_ptr(code_c), _c(codes, np.uint8).reshape(-1), _c(absmax, np.float32), out,
                                      blocksize, n, qt)
    return round_to(out, dtype)


def nested_absmax(absmax2, absmax_8bit, code2, offset: float) -> np.ndarray:
    a8 = _c(absmax_8bit, np.uint8).reshape(-1)
    out = np.empty(a8.size, np.float32)
    lib().oracle_nested_absmax(_c(absmax2, np.float32), a8, _c(code2, np.float32), float(offset), out, a8.size)
    return out


def gemm_4bit(A_f32, B_codes, absmax, M, N, K, blocksize, quant_type, wdtype="bf16", bias_f32=None,
              absmax_8bit=None, absmax_code=None, absmax_offset=None) -> np.ndarray:
    """Double-precision accumulation of A . W_T^T (+bias); returns float64 [M, N] (unrounded)."""
    out = np.empty((M, N), np.float64)
    a8 = _c(absmax_8bit, np.uint8) if absmax_8bit is not None else None
    ac = _c(absmax_code, np.float32) if absmax_code is not None else None
    ao = _c(np.asarray([absmax_offset]), np.float32) if absmax_offset is not None else None
    b = _c(bias_f32, np.float32) if bias_f32 is not None else None
    lib().oracle_gemm_4bit(_c(A_f32, np.float32).reshape(-1), _c(B_codes, np.uint8).reshape(-1),
                           _c(absmax, np.float32), _ptr(a8), _ptr(ac), _ptr(ao), out.reshape(-1), _ptr(b), M, N, K,
                           blocksize, QUANT_TYPE[quant_type], {"fp32": 0, "bf16": 1, "fp16": 2}[wdtype])
    return out


# ----------------------------------------------------------------------------------------
# LLM.int8()
# ----------------------------------------------------------------------------------------
def int8_vector_quant(A_fp16_bits: np.ndarray, threshold: float = 0.0):
    A = _c(A_fp16_bits, np.uint16)
    rows, cols = A.shape
    out = np.empty((rows, cols), np.int8)
    stats = np.empty(rows, np.float32)
    lib().oracle_int8_vector_quant(A.reshape(-1), out.reshape(-1), stats, float(threshold), rows, cols)
    return out, stats


def int8_gemm(A_i8: np.ndarray, B_i8: np.ndarray) -> np.ndarray:
    A = _c(A_i8, np.int8)
    B = _c(B_i8, np.int8)
    M, K = A.shape
    N = B.shape[0]
    C = np.empty((M, N), np.int32)
    lib().oracle_int8_gemm(A.reshape(-1), B.reshape(-1), C.reshape(-1), M, N, K)
    return C


def int8_mm_dequant(A_i32, row_stats, col_stats, bias_fp16_bits=None) -> np.ndarray:
    A = _c(A_i32, np.int32)
    rows, cols = A.shape
    out = np.empty((rows, cols), np.uint16)
    b = _c(bias_fp16_bits, np.uint16) if bias_fp16_bits is not None else None
    lib().oracle_int8_mm_dequant(A.reshape(-1), _c(row_stats, np.float32), _c(col_stats, np.float32), out.reshape(-1),
                                 _ptr(b), rows, cols)
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())


# ----------------------------------------------------------------------------------------
# reference libraries built from the reference sources (oracle/_ref)
# ----------------------------------------------------------------------------------------
def ref_cuda_library_path() -> Path:
    return REF_DIR / "libbitsandbytes_cuda_ref.so"


def ref_cpu_library_path() -> Path | None:
    """The as-shipped (AVX-512) build if this host supports it, else the AVX2 build."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    need = ("avx512f", "avx512bw", "avx512dq", "avx512vl", "avx512_bf16")
    cand = REF_DIR / ("libbitsandbytes_cpu.so" if all(k in flags for k in need) else "libbitsandbytes_cpu_avx2.so")
    return cand if cand.exists() else None
