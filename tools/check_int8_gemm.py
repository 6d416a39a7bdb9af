"""Correctness + timing driver for the int8 GEMM variants.  usage: debug_i8p.py [pair|multicast]"""
import os, sys
os.environ["BNB_B200_I8_MODE"] = sys.argv[1] if len(sys.argv) > 1 else "pair"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat

def run(M, N, K, epi=0):
    print(f"-- M{M} N{N} K{K} epi{epi}", flush=True)
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    CA = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda", generator=g)
    CB = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda", generator=g)
    C = torch.full((M, N), -7, dtype=torch.int32, device="cuda")
    rc = nat.lib.cigemmlt_32(None, N, M, K, CB.data_ptr(), CA.data_ptr(), C.data_ptr(), None, K, K, N, nat.stream())
    torch.cuda.synchronize()
    ref = (CA.double() @ CB.double().t()).to(torch.int32)
    bad = (C != ref)
    print(f"   rc={rc} mismatches={int(bad.sum())} of {M*N}", flush=True)
    if bad.any():
        idx = bad.nonzero()
        print("   first bad", idx[:4].tolist(), "rows bad:", int(bad.any(1).sum()), "cols bad:", int(bad.any(0).sum()), flush=True)
    return not bad.any()

ok = True
for shp in ((256, 256, 128), (256, 256, 1024), (128, 256, 128), (9, 24, 64), (300, 700, 192), (4096, 11008, 4096)):
    ok &= run(*shp)
if ok:
    from tools.probe_perf import timeit
    for (M, K, N) in ((4096, 4096, 11008), (4096, 4096, 4096), (8192, 8192, 8192)):
        CA = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
        CB = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda")
        SCA = torch.rand(M, device="cuda") + 0.5
        SCB = torch.rand(N, device="cuda") + 0.5
        C = torch.empty(M, N, dtype=torch.int32, device="cuda")
        o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
        ops = 2.0 * M * N * K
        t1, _ = timeit(lambda: nat.lib.cigemmlt_32(None, N, M, K, CB.data_ptr(), CA.data_ptr(), C.data_ptr(), None, K, K, N, nat.stream()), iters=10)
        t2, _ = timeit(lambda: nat.lib.cbnb_b200_int8_scaled_mm(CA.data_ptr(), CB.data_ptr(), SCA.data_ptr(), SCB.data_ptr(), None, o16.data_ptr(), M, N, K, 1, nat.stream()), iters=10)
        t3, _ = timeit(lambda: torch._int_mm(CA, CB.t()), iters=10)
        print(f"mode={os.environ['BNB_B200_I8_MODE']} M{M} K{K} N{N}: i32 {t1:.1f} us ({ops/t1/1e6:.0f} TOPS) | fused {t2:.1f} us ({ops/t2/1e6:.0f} TOPS) | cublasLt _int_mm {t3:.1f} us ({ops/t3/1e6:.0f} TOPS)", flush=True)
print("done ok=", ok, flush=True)
