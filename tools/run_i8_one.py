"""One int8 GEMM shape, a few launches (for ncu captures and quick timing).  usage: run_i8_one.py [M N K]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (4096, 11008, 4096)
CA = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
CB = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda")
SCA = torch.rand(M, device="cuda") + 0.5
SCB = torch.rand(N, device="cuda") + 0.5
o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
def f():
    nat.lib.cbnb_b200_int8_scaled_mm(CA.data_ptr(), CB.data_ptr(), SCA.data_ptr(), SCB.data_ptr(), None, o16.data_ptr(), M, N, K, 1, nat.stream())
for _ in range(4):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    f()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"mode={os.environ.get('BNB_B200_I8_MODE','pair')} M{M} N{N} K{K}: fused {us:.1f} us ({2.0*M*N*K/us/1e6:.0f} TOPS)", flush=True)
nat.check()
