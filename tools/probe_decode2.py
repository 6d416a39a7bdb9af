"""Decode-regime (M <= 16) kernels, graph-replay timed over rotating weight copies larger than L2:
forced path 0 (CUDA-core GEMV), 3 (mma.sync decode kernel), 1 (tcgen05).  BNB_B200_GEMV_WARPS selects the
warps per CTA of path 3 (static: one process per setting)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from benchmarks.common import time_us  # noqa: E402
from tests import _native as nat  # noqa: E402
from tests.test_gpu_gemm4 import make_problem  # noqa: E402
from tools.probe_perf import run_nosync  # noqa: E402

shapes = ((4096, 4096), (14336, 4096), (4096, 14336), (1024, 4096))
print("BNB_B200_GEMV_WARPS =", os.environ.get("BNB_B200_GEMV_WARPS"))
for (N, K) in shapes:
    sets = max(3, int(160e6 // (N * K // 2)) + 1)  # weight copies: beyond the 126 MB L2
    for M in (1, 4, 8, 16):
        p0 = make_problem(M, N, K, "nf4", "bf16")
        p0.pop("_out", None)
        copies = [dict(p0, packed=p0["packed"].clone(), absmax=p0["absmax"].clone()) for _ in range(sets)]
        res = []
        for path in (0, 3, 1):
            if path == 0 and M > 4:
                continue
            nat.lib.cbnb_b200_gemm_4bit_force_path(path)
            us, mode = time_us(lambda i: run_nosync(nat.lib, copies[i % sets]), 40)
            nat.lib.cbnb_b200_gemm_4bit_force_path(-1)
            gb = N * K / 2 / us / 1e3
            res.append(f"path{path} {us:.1f} us ({gb:.0f} GB/s)")
        print(f"M{M} N{N} K{K}: " + " | ".join(res), flush=True)
nat.check()
print("done")
