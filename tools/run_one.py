"""Run one kernel configuration a few times (for ncu captures).  usage:
   python tools/run_one.py gemm M N K [path] | dequant | quant8 | quant4 | int8"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat  # noqa: E402
from tests.test_gpu_gemm4 import make_problem  # noqa: E402
from tools.probe_perf import run_nosync  # noqa: E402

what = sys.argv[1]
iters = 3
if what == "gemm":
    M, N, K = (int(v) for v in sys.argv[2:5])
    path = int(sys.argv[5]) if len(sys.argv) > 5 else -1
    p = make_problem(M, N, K, "nf4", "bf16")
    nat.lib.cbnb_b200_gemm_4bit_force_path(path)
    for _ in range(iters):
        run_nosync(nat.lib, p)
    torch.cuda.synchronize()
elif what in ("dequant", "quant8", "quant4"):
    from bitsandbytes_b200.functional import create_dynamic_map

    code = create_dynamic_map().cuda()
    if what == "quant8":
        n = 4 * 1024 * 1024
        A = torch.randn(n, device="cuda")
        absmax = torch.empty(n // 4096, device="cuda")
        q = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(iters):
            nat.lib.cbnb_b200_quantize_blockwise(code.data_ptr(), A.data_ptr(), absmax.data_ptr(), q.data_ptr(), 4096, n, 0, 0, nat.stream())
    else:
        n = 4096 * 4096
        W = torch.randn(n, device="cuda", dtype=torch.bfloat16)
        absmax = torch.empty(n // 64, device="cuda")
        q4 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
        out = torch.empty(n, device="cuda", dtype=torch.bfloat16)
        for _ in range(iters):
            nat.lib.cbnb_b200_quantize_blockwise(None, W.data_ptr(), absmax.data_ptr(), q4.data_ptr(), 64, n, 2, 2, nat.stream())
            if what == "dequant":
                nat.lib.cdequantize_blockwise_bf16_nf4(None, q4.data_ptr(), absmax.data_ptr(), out.data_ptr(), 64, n, nat.stream())
    torch.cuda.synchronize()
elif what == "int8":
    M, K, N = 4096, 4096, 11008
    CA = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    CB = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda")
    SCA = torch.rand(M, device="cuda") + 0.5
    SCB = torch.rand(N, device="cuda") + 0.5
    o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
    for _ in range(iters):
        nat.lib.cbnb_b200_int8_scaled_mm(CA.data_ptr(), CB.data_ptr(), SCA.data_ptr(), SCB.data_ptr(), None, o16.data_ptr(), M, N, K, 1, nat.stream())
    torch.cuda.synchronize()
nat.check()
print("done")
