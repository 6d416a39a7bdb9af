"""A/B of the cta_group::2 pair variant of the 4-bit GEMM (BNB_B200_PAIR, read per call):
bit-equality with the default kernel, then timing.  usage: probe_pair.py [MxNxK ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem, run
from tools.probe_perf import run_nosync, timeit

def with_pair(v, fn):
    os.environ["BNB_B200_PAIR"] = str(v)
    try:
        return fn()
    finally:
        os.environ["BNB_B200_PAIR"] = "0"

ok = True
for (M, N, K, qt, dt, kw) in ((256, 256, 128, "nf4", "bf16", {}), (300, 512, 320, "fp4", "fp16", dict(bias=True)),
                              (257, 384, 64, "nf4", "bf16", dict(nested=True)), (1000, 1024, 1024, "nf4", "bf16", {}),
                              (4096, 4096, 4096, "nf4", "bf16", {})):
    p = make_problem(M, N, K, qt, dt, **kw)
    a = with_pair(0, lambda: run(nat.lib, p)).clone()
    b = with_pair(1, lambda: run(nat.lib, p)).clone()
    torch.cuda.synchronize()
    nat.check()
    bad = int((a.view(torch.int16) != b.view(torch.int16)).sum())
    print(f"eq {M}x{N}x{K} {qt} {dt} {kw}: {'OK' if bad == 0 else 'MISMATCH'} ({bad} of {a.numel()})", flush=True)
    ok &= bad == 0
shapes = [(4096, 4096, 4096), (4096, 11008, 4096), (2048, 14336, 4096), (1024, 4096, 4096), (8192, 8192, 8192)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    p = make_problem(M, N, K, "nf4", "bf16")
    t0, m0 = with_pair(0, lambda: timeit(lambda: run_nosync(nat.lib, p), iters=15))
    t1, m1 = with_pair(1, lambda: timeit(lambda: run_nosync(nat.lib, p), iters=15))
    fl = 2.0 * M * N * K
    print(f"{M}x{N}x{K}: default {t0:.1f} us ({fl/t0/1e6:.0f} TF, min {m0:.1f}) | pair {t1:.1f} us ({fl/t1/1e6:.0f} TF, min {m1:.1f})", flush=True)
nat.check()
print("done ok=", ok)
