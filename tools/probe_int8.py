import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tools.probe_perf import timeit
for (M, K, N) in ((4096, 4096, 11008), (4096, 4096, 4096), (256, 4096, 11008), (8192, 8192, 8192)):
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
    def chain():
        c = torch._int_mm(CA, CB.t())
        nat.lib.cdequant_mm_int32_fp16(c.data_ptr(), SCA.data_ptr(), SCB.data_ptr(), o16.data_ptr(), None, M, N, nat.stream())
    t4, _ = timeit(chain, iters=10)
    print(f"M{M} K{K} N{N}: ours i32 {t1:.1f} us ({ops/t1/1e6:.0f} TOPS) | ours fused {t2:.1f} us ({ops/t2/1e6:.0f} TOPS) | cublasLt _int_mm {t3:.1f} us ({ops/t3/1e6:.0f} TOPS) | _int_mm + dequant kernel {t4:.1f} us", flush=True)
nat.check()
