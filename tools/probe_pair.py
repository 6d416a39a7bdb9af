"""A/B of the experimental variants of the 4-bit GEMM (switches read per call): the cta_group::2 pair
(BNB_B200_PAIR), all 16 decode warps per stage (BNB_B200_DECODE16) and both -- bit-equality with the
default kernel first, then timing.  usage: probe_pair.py [MxNxK ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem, run
from tools.probe_perf import run_nosync, timeit

VARIANTS = {"default": (0, 0), "pair": (1, 0), "decode16": (0, 1), "pair+decode16": (1, 1)}


def with_variant(name, fn):
    pair, d16 = VARIANTS[name]
    os.environ["BNB_B200_PAIR"], os.environ["BNB_B200_DECODE16"] = str(pair), str(d16)
    try:
        return fn()
    finally:
        os.environ["BNB_B200_PAIR"] = os.environ["BNB_B200_DECODE16"] = "0"

ok = True
for (M, N, K, qt, dt, kw) in ((256, 256, 128, "nf4", "bf16", {}), (300, 512, 320, "fp4", "fp16", dict(bias=True)),
                              (257, 384, 64, "nf4", "bf16", dict(nested=True)), (1000, 1024, 1024, "nf4", "bf16", {}),
                              (4096, 4096, 4096, "nf4", "bf16", {})):
    p = make_problem(M, N, K, qt, dt, **kw)
    a = with_variant("default", lambda: run(nat.lib, p)).clone()
    for v in ("pair", "decode16", "pair+decode16"):
        b = with_variant(v, lambda: run(nat.lib, p)).clone()
        torch.cuda.synchronize()
        nat.check()
        bad = int((a.view(torch.int16) != b.view(torch.int16)).sum())
        print(f"eq {v} {M}x{N}x{K} {qt} {dt} {kw}: {'OK' if bad == 0 else 'MISMATCH'} ({bad} of {a.numel()})", flush=True)
        ok &= bad == 0
shapes = [(4096, 4096, 4096), (4096, 11008, 4096), (2048, 14336, 4096), (1024, 4096, 4096), (8192, 8192, 8192)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    p = make_problem(M, N, K, "nf4", "bf16")
    fl = 2.0 * M * N * K
    parts = []
    for v in VARIANTS:
        t, m = with_variant(v, lambda: timeit(lambda: run_nosync(nat.lib, p), iters=15))
        parts.append(f"{v} {t:.1f} us ({fl/t/1e6:.0f} TF)")
    print(f"{M}x{N}x{K}: " + " | ".join(parts), flush=True)
nat.check()
print("done ok=", ok)
