import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync
for (M, N, K) in ((1, 4096, 4096), (4, 4096, 4096), (8, 4096, 4096), (16, 4096, 4096), (64, 4096, 4096), (256, 4096, 4096), (1, 14336, 4096), (1, 4096, 14336)):
    p = make_problem(M, N, K, "nf4", "bf16")
    for path in (0, 1):
        if path == 0 and M > 8: continue
        nat.lib.cbnb_b200_gemm_4bit_force_path(path)
        for _ in range(3): run_nosync(nat.lib, p)
        torch.cuda.synchronize()
nat.check()
