"""One 4-bit GEMM shape, a few launches (for ncu captures).  usage: run_gemm4_one.py M N K"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync
M, N, K = (int(v) for v in sys.argv[1:4])
p = make_problem(M, N, K, "nf4", "bf16")
for _ in range(4):
    run_nosync(nat.lib, p)
torch.cuda.synchronize()
nat.check()
