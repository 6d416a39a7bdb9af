"""Round-2 starting point: the persistent variant of the 4-bit GEMM (BNB_B200_PERSISTENT=1, written
but never run in round 1) against the default kernel -- bit equality first, then timing.
Each variant runs in its own process because the switch is read once per process.
usage: python tools/check_persistent.py"""
import json, os, subprocess, sys

code = r'''
import os, sys, json, hashlib, torch
sys.path.insert(0, os.getcwd())
from tests import _native as nat
from tests.test_gpu_gemm4 import make_problem, run
from tools.probe_perf import run_nosync, timeit
res = {}
for (M, N, K, kw) in ((4096, 4096, 4096, {}), (2500, 4096, 1088, dict(bias=True)), (4096, 11008, 4096, {}),
                      (8192, 8192, 8192, {})):
    p = make_problem(M, N, K, "nf4", "bf16", **kw)
    out = run(nat.lib, p)
    torch.cuda.synchronize()
    nat.check()
    h = hashlib.sha1(out.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:16]
    t, t0 = timeit(lambda: run_nosync(nat.lib, p), iters=15)
    res[f"{M}x{N}x{K}"] = {"sha1": h, "us": round(t, 1), "min_us": round(t0, 1)}
print(json.dumps(res))
'''
out = {}
for mode in ("0", "1"):
    env = dict(os.environ, BNB_B200_PERSISTENT=mode)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else None
    if line is None:
        print(f"mode {mode} failed:\n{r.stderr[-2000:]}")
        sys.exit(1)
    out[mode] = json.loads(line)
ok = True
for shape in out["0"]:
    a, b = out["0"][shape], out["1"][shape]
    same = a["sha1"] == b["sha1"]
    ok &= same
    print(f"{shape}: default {a['us']} us | persistent {b['us']} us | {'bit-identical' if same else 'MISMATCH'}")
sys.exit(0 if ok else 2)
