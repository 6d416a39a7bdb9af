"""Decode-regime (M <= 16) kernels side by side: forced path 0 (CUDA-core GEMV), 3 (mma.sync decode
kernel), 1 (tcgen05).  Run under `ncu --metrics gpu__time_duration.sum` for true kernel times
(event pairs around ~10 us kernels are dominated by launch latency); prints event times otherwise."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync, timeit

quick = len(sys.argv) > 1 and sys.argv[1] in ("once", "once3")
only3 = len(sys.argv) > 1 and sys.argv[1] == "once3"
for (N, K) in ((4096, 4096), (14336, 4096), (4096, 14336)):
    for M in (1, 2, 4, 8, 12, 16):
        p = make_problem(M, N, K, "nf4", "bf16")
        res = []
        for path in (0, 3, 1):
            if (path == 0 and M > 4) or (only3 and path != 3):
                continue
            nat.lib.cbnb_b200_gemm_4bit_force_path(path)
            if quick:
                run_nosync(nat.lib, p)
                torch.cuda.synchronize()
            else:
                t, t0 = timeit(lambda: run_nosync(nat.lib, p), iters=20)
                res.append(f"path{path} {t:.1f} (min {t0:.1f})")
            nat.lib.cbnb_b200_gemm_4bit_force_path(-1)
        if not quick:
            print(f"M{M} N{N} K{K}: " + " | ".join(res), flush=True)
nat.check()
print("done")
