import json, os, subprocess, sys
code = r'''
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem
from tools.probe_perf import run_nosync, timeit
res = {}
for (M, N, K) in ((4096, 4096, 4096), (1024, 4096, 4096)):
    p = make_problem(M, N, K, "nf4", "bf16")
    t, t0 = timeit(lambda: run_nosync(nat.lib, p), iters=10)
    res[f"{M}x{N}x{K}"] = round(t, 1)
print(json.dumps(res))
'''
for dbg in (sys.argv[1:] or ("0", "1", "2", "3", "4", "8", "7", "15", "12")):
    env = dict(os.environ, BNB_B200_DEBUG=dbg, BNB_B200_CLUSTER="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("debug", dbg, out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:], flush=True)
