"""One decode-regime launch per path for ncu captures.  usage: run_decode_one.py M N K path"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync
M, N, K, path = (int(v) for v in sys.argv[1:5])
p = make_problem(M, N, K, "nf4", "bf16")
nat.lib.cbnb_b200_gemm_4bit_force_path(path)
for _ in range(3):
    run_nosync(nat.lib, p)
torch.cuda.synchronize()
nat.check()
