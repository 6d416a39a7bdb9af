"""probe: M=4096 GEMMs under BNB_B200_CLUSTER={1,2,4} (each in a fresh process)."""
import json, os, subprocess, sys
code = r'''
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync, timeit
res = {}
for (M, N, K) in ((4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008), (1024, 4096, 4096), (256, 4096, 4096), (2048, 14336, 4096)):
    p = make_problem(M, N, K, "nf4", "bf16")
    t, t0 = timeit(lambda: run_nosync(nat.lib, p), iters=15)
    res[f"{M}x{N}x{K}"] = (round(t, 1), round(2.0 * M * N * K / t / 1e6, 1))
nat.check()
print(json.dumps(res))
'''
for cl in ("1", "2", "4", "-1"):
    env = dict(os.environ, BNB_B200_CLUSTER=cl)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("cluster", cl, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-800:], flush=True)
