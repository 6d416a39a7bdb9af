"""Developer probe for the CTA-pair 4-bit GEMM (csrc/gemm4_pair.cu): bit-equality with the one-CTA tcgen05
kernel on ragged / nested / split shapes, timing per token tile, and an event trace of cluster 0.
usage: probe_gemm4_pair.py [eq] [time] [trace] [MxNxK ...]"""
import os
import sys

os.environ["BNB_B200_PAIR_KERNEL"] = "0"  # cgemm_4bit_* = the one-CTA kernel; the pair kernel through its own entry
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import _native as nat  # noqa: E402
from tests.test_gpu_gemm4 import make_problem, run  # noqa: E402
from tools.probe_perf import timeit  # noqa: E402

L = nat.lib


def run_pair(p, mt=0, splits=0, trace=None, out=None, sync=True):
    M, N, K = p["M"], p["N"], p["K"]
    if out is None:
        out = torch.full((M, N), float("nan"), device="cuda", dtype=nat.DTYPE[p["dtype"]])
    rc = L.cbnb_b200_gemm_4bit_pair(nat.ptr(p["x"]), nat.ptr(p["packed"]), nat.ptr(p["absmax"]), nat.ptr(p["absmax_8bit"]),
                                    nat.ptr(p["absmax_code"]), nat.ptr(p["absmax_offset"]), nat.ptr(out), nat.ptr(p["bias"]),
                                    M, N, K, N, p["bs"], nat.QT_ID[p["qt"]], nat.DTYPE_ID[p["dtype"]], mt, splits,
                                    nat.ptr(trace), nat.stream())
    if sync:
        torch.cuda.synchronize()
    return out if rc == 0 else None


def run_old_nosync(p, out):
    fn = getattr(L, f"cgemm_4bit_{p['dtype']}")
    fn(nat.ptr(p["x"]), nat.ptr(p["packed"]), nat.ptr(p["absmax"]), nat.ptr(p["absmax_8bit"]), nat.ptr(p["absmax_code"]),
       nat.ptr(p["absmax_offset"]), nat.ptr(out), nat.ptr(p["bias"]), p["M"], p["N"], p["K"], p["bs"], nat.QT_ID[p["qt"]],
       nat.stream())


def eq():
    ok = True
    cases = [
        (512, 256, 128, "nf4", "bf16", {}),
        (600, 512, 320, "fp4", "fp16", dict(bias=True)),
        (513, 384, 192, "nf4", "bf16", dict(nested=True)),
        (1000, 1024, 1024, "nf4", "bf16", {}),
        (777, 1000, 704, "nf4", "bf16", dict(bs=32, bias=True)),
        (640, 768, 512, "fp4", "bf16", dict(bs=128, nested=True)),
        (4096, 4096, 4096, "nf4", "bf16", {}),
    ]
    for (M, N, K, qt, dt, kw) in cases:
        p = make_problem(M, N, K, qt, dt, **kw)
        a = run(L, p)
        nat.check()
        for mt in (128, 256, 384):
            for sp in (0, 2, 102):
                b = run_pair(p, mt, sp)
                nat.check()
                if b is None:
                    # forcing a split of EVERY tile of a large problem exceeds the fixed split-K workspace
                    print(f"eq {M}x{N}x{K} {qt} {dt} {kw} mt={mt} splits={sp}: not served"
                          f"{'' if sp == 2 else ' (UNEXPECTED)'}", flush=True)
                    ok &= sp == 2
                    continue
                bad = int((a.view(torch.int16) != b.view(torch.int16)).sum())
                # split K changes the fp32 summation order: compare within a few ulp instead of bit-equal
                # (sp = 0 is the production rule, which splits the partial last wave of large problems)
                if bad:  # (sp = 0 is the production rule, which may split the partial last round)
                    rel = float((a.float() - b.float()).norm() / a.float().norm())
                    good = rel < 2e-3 and not torch.isnan(b.float()).any()
                    print(f"eq {M}x{N}x{K} {qt} {dt} {kw} mt={mt} splits={sp}: {'ok' if good else 'MISMATCH'} "
                          f"(rel {rel:.2e}, {bad} differ)", flush=True)
                    ok &= good
                else:
                    print(f"eq {M}x{N}x{K} {qt} {dt} {kw} mt={mt} splits={sp}: {'OK' if bad == 0 else 'MISMATCH'} "
                          f"({bad} of {a.numel()})", flush=True)
                    ok &= bad == 0
        # determinism of the split path
        b1 = run_pair(p, 256, 102)
        b2 = run_pair(p, 256, 102)
        same = b1 is not None and bool((b1.view(torch.int16) == b2.view(torch.int16)).all())
        print(f"   split determinism: {'OK' if same else 'MISMATCH'}", flush=True)
        ok &= same
    print("eq done ok=", ok, flush=True)
    return ok


def time_shapes(shapes):
    for (M, N, K) in shapes:
        p = make_problem(M, N, K, "nf4", "bf16")
        fl = 2.0 * M * N * K
        out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        parts = []
        t, m = timeit(lambda: run_old_nosync(p, out), iters=15)
        parts.append(f"one-CTA {t:.1f} us ({fl/t/1e6:.0f} TF)")
        for mt, sp in ((256, 0), (256, 1), (384, 0), (384, 1)):
            t, m = timeit(lambda: run_pair(p, mt, sp, out=out, sync=False), iters=15)
            parts.append(f"pair mt={mt} sp={sp} {t:.1f} us ({fl/t/1e6:.0f} TF, min {m:.1f})")
        # cuBLAS bf16 for context
        W = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
        t, m = timeit(lambda: torch.matmul(p["x"], W.t(), out=out), iters=15)
        parts.append(f"cuBLAS bf16 {t:.1f} us ({fl/t/1e6:.0f} TF)")
        print(f"{M}x{N}x{K}: " + " | ".join(parts), flush=True)
        nat.check()


ROLES = ["x_issue", "mma_xfull", "mma_issued", "dec_cfull", "dec_math", "dec_empty", "dec_arrived", "mma_afull", "epi_begin", "epi_end"]


def trace_lite(shape, mt, sp=1):
    """Only four landmarks per tile (no per-stage stores): the undisturbed duration of the main loop and the epilogue."""
    M, N, K = shape
    p = make_problem(M, N, K, "nf4", "bf16")
    os.environ["BNB_B200_TRACE_LITE"] = "1"
    tr = torch.zeros(2 * 10 * 256 + 4 * 1024, dtype=torch.int64, device="cuda")
    run_pair(p, mt, sp)
    run_pair(p, mt, sp, trace=tr)
    os.environ["BNB_B200_TRACE_LITE"] = "0"
    nat.check()
    full = tr.cpu().numpy()
    t = full[:2 * 10 * 256].reshape(2, 10, 256)
    tl = full[2 * 10 * 256:].reshape(512, 8)
    tl = tl[tl[:, 0] > 0]
    if len(tl):
        t0 = tl[:, 0].min()
        st, en, pro = tl[:, 0] - t0, tl[:, 1] - t0, tl[:, 3]
        items = tl[:, 4:8] - t0
        print(f"timeline {M}x{N}x{K} mt={mt} sp={sp}: {len(tl)} persistent clusters; kernel span {en.max()} ns; first start spread "
              f"{st.max()} ns; entry -> first MMA median {np.median(pro):.0f} cycles (p90 {np.percentile(pro,90):.0f})")
        prev = st
        for k in range(4):
            ids = items[:, k] > 0
            if not ids.any():
                break
            e = items[ids, k]
            d = e - prev[ids]
            print(f"   item {k}: {int(ids.sum())} clusters, end median {np.median(e):.0f} ns (max {e.max()}), duration median "
                  f"{np.median(d):.0f} ns (p10 {np.percentile(d,10):.0f}, p90 {np.percentile(d,90):.0f})")
            prev = np.where(ids, items[:, k], prev)
        print(f"   cluster end median {np.median(en):.0f} ns, max {en.max()} ns")
    first, last_issued, acc, epi_end = t[0][1][0], t[0][2].max(), t[0][8][0], t[0][9][0]
    ns = t[0][9][41] - t[0][9][40]
    cyc = epi_end - first
    nst = (K // 64) // (2 if sp == 2 else 1)
    print(f"lite {M}x{N}x{K} mt={mt} sp={sp}: main loop {last_issued - first} cycles = {(last_issued - first) / max(nst, 1):.0f} per a-stage "
          f"({nst} stages); last issue -> accumulator ready {acc - last_issued}; epilogue {epi_end - acc}; "
          f"{cyc} cycles in {ns} ns = {cyc / max(ns, 1):.3f} GHz", flush=True)


def trace(shape, mt):
    M, N, K = shape
    p = make_problem(M, N, K, "nf4", "bf16")
    tr = torch.zeros(2 * 10 * 256, dtype=torch.int64, device="cuda")
    run_pair(p, mt, 1)  # warm
    run_pair(p, mt, 1, trace=tr)
    nat.check()
    t = tr.cpu().numpy().reshape(2, 10, 256)
    nst = min(256, K // 64)
    for cta in (0, 1):
        base = t[cta][t[cta] > 0].min()
        print(f"--- trace {M}x{N}x{K} mt={mt} cta {cta} (cycles since first event; per a-stage)")
        print("  i " + " ".join(f"{r:>11s}" for r in ROLES[:8]))
        for i in list(range(0, min(nst, 20))) + list(range(max(20, nst - 4), nst)):
            print(f"{i:3d} " + " ".join(f"{(t[cta][r][i] - base) if t[cta][r][i] else -1:11d}" for r in range(8)))
        print(f"  epilogue begin {t[cta][8][0]-base}, end {t[cta][9][0]-base}")
        if cta == 0:
            iss = (t[0][2][4:nst] - t[0][1][4:nst]).astype(np.int64)      # MMAs + probes + commits of a stage
            tail = (t[0][7][4:nst] - t[0][2][4:nst]).astype(np.int64)     # fallback waits after a failed probe
            xl = (t[0][1][4:nst] - t[0][0][4:nst]).astype(np.int64)       # activation load: issue -> stage start
            print(f"  MMA thread: issue block {np.median(iss):.0f} (p90 {np.percentile(iss,90):.0f}), fallback waits {np.median(tail):.0f} "
                  f"(p90 {np.percentile(tail,90):.0f}); x issue->stage start {np.median(xl):.0f} (p90 {np.percentile(xl,90):.0f})")
            for st in (0, 1):
                mt_ = t[0][9][1 + 16 * st: 1 + 16 * st + 12].astype(np.int64)
                mt_ = mt_[mt_ > 0]
                if len(mt_) > 1:
                    print(f"  a-stage {24 + st} issue deltas (after each tcgen05.mma, then the two commits): " + " ".join(str(int(v)) for v in np.diff(mt_)))
        dm = (t[cta][4][:nst] - t[cta][3][:nst]).astype(np.int64)
        de = (t[cta][5][:nst] - t[cta][4][:nst]).astype(np.int64)
        da = (t[cta][6][:nst] - t[cta][5][:nst]).astype(np.int64)
        print(f"  decode: c_full->math {np.median(dm):.0f} (p90 {np.percentile(dm,90):.0f}); math->empty seen {np.median(de):.0f} (p90 {np.percentile(de,90):.0f}); st+arrive {np.median(da):.0f}")


if __name__ == "__main__":
    args = sys.argv[1:]
    shapes = [tuple(int(v) for v in a.split("x")) for a in args if "x" in a]
    ok = True
    if not args or "eq" in args:
        ok = eq()
    if not args or "time" in args:
        time_shapes(shapes or [(4096, 4096, 4096), (4096, 11008, 4096), (4096, 4096, 11008), (2048, 14336, 4096),
                               (1024, 4096, 4096), (512, 4096, 4096), (8192, 8192, 8192)])
    if not args or "lite" in args:
        for mt in (128, 256, 384):
            trace_lite((shapes or [(4096, 4096, 4096)])[0], mt, 1)
        # a split tile (every tile split two ways; small enough for the fixed workspace)
        for mt in (256, 384):
            trace_lite((1024, 4096, 4096), mt, 1)
            trace_lite((1024, 4096, 4096), mt, 2)
    if not args or "trace" in args:
        for mt in (256, 384):
            trace((shapes or [(4096, 4096, 4096)])[0], mt)
    print("done ok=", ok)
