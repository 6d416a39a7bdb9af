"""Quick in-process timing of the fused GEMM on a few shapes (one torch import)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync, timeit
shapes = [(4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008), (1024, 4096, 4096), (256, 4096, 4096),
          (64, 4096, 4096), (16, 4096, 4096), (1, 4096, 4096), (16, 11008, 4096), (2048, 14336, 4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    p = make_problem(M, N, K, "nf4", "bf16")
    t, t0 = timeit(lambda: run_nosync(nat.lib, p), iters=15)
    extra = ""
    if M <= 16:
        nat.lib.cbnb_b200_gemm_4bit_force_path(0)
        ts, ts0 = timeit(lambda: run_nosync(nat.lib, p), iters=15)
        nat.lib.cbnb_b200_gemm_4bit_force_path(1)
        tt, tt0 = timeit(lambda: run_nosync(nat.lib, p), iters=15)
        nat.lib.cbnb_b200_gemm_4bit_force_path(-1)
        tn, _ = timeit(lambda: run_nosync(nat.lib, p), iters=15, flush=False)
        extra = f" | simt {ts:.1f} (min {ts0:.1f})  tc {tt:.1f} (min {tt0:.1f})  auto-noflush {tn:.1f}"
    print(f"{M}x{N}x{K}: {t:.1f} us  {2.0 * M * N * K / t / 1e6:.1f} TFLOPS (min {t0:.1f}){extra}", flush=True)
nat.check()
