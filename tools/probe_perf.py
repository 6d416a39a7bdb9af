"""Developer probe (not part of the product or the bench contract): times our kernels, the
reference CUDA library's kernels (oracle/_ref, same C ABI) and cuBLAS on the BASELINE shapes.
Writes one JSON object per line to gpurun_out/probe.jsonl."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _native as nat  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
flush_buf = None


def timeit(fn, iters=20, warm=3, flush=True):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.view(torch.int32).sum()  # read-only sweep > L2: evicts, leaves no dirty lines
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(f, **kw):
    print(json.dumps(kw), flush=True)
    f.write(json.dumps(kw) + "\n")
    f.flush()


def main():
    ref = nat.ref_cuda()
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    with open(os.path.join(OUT, "probe.jsonl"), "a") as f:
        emit(f, what="env", gpu=torch.cuda.get_device_name(0), ref_cuda=ref is not None, only=only)
        from bitsandbytes_b200.functional import create_dynamic_map

        code = create_dynamic_map().cuda()
        if only in ("all", "blockwise"):
            n = 4 * 1024 * 1024
            A = torch.randn(n, device="cuda")
            for bs in (4096, 256):
                absmax = torch.empty(n // bs, device="cuda")
                q = torch.empty(n, dtype=torch.uint8, device="cuda")
                out = torch.empty(n, device="cuda")
                for name, L in (("ours", nat.lib), ("ref", ref)):
                    if L is None:
                        continue
                    if name == "ours":
                        fq = lambda: L.cbnb_b200_quantize_blockwise(code.data_ptr(), A.data_ptr(), absmax.data_ptr(), q.data_ptr(), bs, n, 0, 0, nat.stream())
                    else:
                        fq = lambda: L.cquantize_blockwise_fp32(code.data_ptr(), A.data_ptr(), absmax.data_ptr(), q.data_ptr(), bs, n)
                    fd = lambda: L.cdequantize_blockwise_fp32(code.data_ptr(), q.data_ptr(), absmax.data_ptr(), out.data_ptr(), bs, n, nat.stream())
                    tq, tq0 = timeit(fq)
                    td, td0 = timeit(fd)
                    byt = 4 * n + n + 4 * n / bs
                    emit(f, what="C1_8bit", impl=name, bs=bs, quant_us=tq, quant_GBs=byt / tq / 1e3, dequant_us=td,
                         dequant_GBs=byt / td / 1e3, quant_min_us=tq0, dequant_min_us=td0)
            # 4-bit weight dequant / quant, 4096x4096 bf16 bs64
            N = K = 4096
            W = torch.randn(N * K, device="cuda", dtype=torch.bfloat16)
            absmax = torch.empty(N * K // 64, device="cuda")
            q4 = torch.empty(N * K // 2, dtype=torch.uint8, device="cuda")
            out = torch.empty(N * K, device="cuda", dtype=torch.bfloat16)
            for name, L in (("ours", nat.lib), ("ref", ref)):
                if L is None:
                    continue
                if name == "ours":
                    fq = lambda: L.cbnb_b200_quantize_blockwise(None, W.data_ptr(), absmax.data_ptr(), q4.data_ptr(), 64, N * K, 2, 2, nat.stream())
                else:
                    fq = lambda: L.cquantize_blockwise_bf16_nf4(None, W.data_ptr(), absmax.data_ptr(), q4.data_ptr(), 64, N * K)
                fd = lambda: L.cdequantize_blockwise_bf16_nf4(None, q4.data_ptr(), absmax.data_ptr(), out.data_ptr(), 64, N * K, nat.stream())
                tq, _ = timeit(fq)
                td, _ = timeit(fd)
                byt = 2 * N * K + N * K / 2 + 4 * N * K / 64
                emit(f, what="C2_nf4_weight", impl=name, quant_us=tq, quant_GBs=byt / tq / 1e3, dequant_us=td,
                     dequant_GBs=byt / td / 1e3)
        if only in ("all", "gemm"):
            from tests.test_gpu_gemm4 import make_problem, run

            for (N, K) in ((4096, 4096), (11008, 4096), (4096, 11008)):
                for M in (1, 4, 16, 64, 256, 1024, 4096):
                    p = make_problem(M, N, K, "nf4", "bf16")
                    flops = 2.0 * M * N * K
                    row = dict(what="C2_gemm", M=M, N=N, K=K)
                    for path, tag in ((1, "tc"), (0, "simt")):
                        if tag == "simt" and M > 16:
                            continue
                        nat.lib.cbnb_b200_gemm_4bit_force_path(path)
                        t, t0 = timeit(lambda: run_nosync(nat.lib, p), iters=10)
                        nat.lib.cbnb_b200_gemm_4bit_force_path(-1)
                        row[f"{tag}_us"] = t
                        row[f"{tag}_TFLOPS"] = flops / t / 1e6
                    if ref is not None:
                        # what the reference does on B200 for M > 4: dequantize + cuBLAS (ops.py:617-623)
                        Wd = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)

                        def ref_fallback():
                            ref.cdequantize_blockwise_bf16_nf4(None, p["packed"].data_ptr(), p["absmax"].data_ptr(), Wd.data_ptr(), 64, N * K, nat.stream())
                            return torch.nn.functional.linear(p["x"], Wd)

                        t, _ = timeit(ref_fallback, iters=10)
                        row["ref_dequant_cublas_us"] = t
                        row["ref_dequant_cublas_TFLOPS"] = flops / t / 1e6
                        if M <= 32:
                            t, _ = timeit(lambda: run_nosync(ref, p), iters=10)
                            row["ref_fused_us"] = t
                    Wd = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
                    t, _ = timeit(lambda: torch.nn.functional.linear(p["x"], Wd), iters=10)
                    row["cublas_bf16_us"] = t
                    row["cublas_bf16_TFLOPS"] = flops / t / 1e6
                    emit(f, **row)
        if only in ("all", "int8"):
            M, K, N = 4096, 4096, 11008
            CA = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
            CB = torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda")
            SCA = torch.rand(M, device="cuda") + 0.5
            SCB = torch.rand(N, device="cuda") + 0.5
            C = torch.empty(M, N, dtype=torch.int32, device="cuda")
            o16 = torch.empty(M, N, dtype=torch.float16, device="cuda")
            ops = 2.0 * M * N * K
            t, _ = timeit(lambda: nat.lib.cigemmlt_32(None, N, M, K, CB.data_ptr(), CA.data_ptr(), C.data_ptr(), None, K, K, N, nat.stream()), iters=10)
            row = dict(what="C3_int8", ours_i32_us=t, ours_i32_TOPS=ops / t / 1e6)
            t, _ = timeit(lambda: nat.lib.cbnb_b200_int8_scaled_mm(CA.data_ptr(), CB.data_ptr(), SCA.data_ptr(), SCB.data_ptr(), None, o16.data_ptr(), M, N, K, 1, nat.stream()), iters=10)
            row.update(ours_fused_us=t, ours_fused_TOPS=ops / t / 1e6)
            if ref is not None:
                ctx = ref.get_context()

                def ref_chain():
                    ref.cigemmlt_32(ctx, N, M, K, CB.data_ptr(), CA.data_ptr(), C.data_ptr(), None, K, K, N, nat.stream())
                    ref.cdequant_mm_int32_fp16(C.data_ptr(), SCA.data_ptr(), SCB.data_ptr(), o16.data_ptr(), None, M, N, nat.stream())

                t, _ = timeit(ref_chain, iters=10)
                row.update(ref_chain_us=t, ref_chain_TOPS=ops / t / 1e6)
            A16 = torch.randn(M, K, device="cuda", dtype=torch.float16)
            q = torch.empty(M, K, dtype=torch.int8, device="cuda")
            st = torch.empty(M, device="cuda")
            t, _ = timeit(lambda: nat.lib.cint8_vector_quant(A16.data_ptr(), q.data_ptr(), st.data_ptr(), 6.0, M, K, nat.stream()))
            row.update(ours_vq_us=t, ours_vq_GBs=(2 * M * K + M * K) / t / 1e3)
            if ref is not None:
                t, _ = timeit(lambda: ref.cint8_vector_quant(A16.data_ptr(), q.data_ptr(), st.data_ptr(), 6.0, M, K, nat.stream()))
                row.update(ref_vq_us=t)
            emit(f, **row)


def run_nosync(L, p):
    out = p.setdefault("_out", torch.empty(p["M"], p["N"], device="cuda", dtype=nat.DTYPE[p["dtype"]]))
    fn = getattr(L, f"cgemm_4bit_{p['dtype']}")
    fn(nat.ptr(p["x"]), nat.ptr(p["packed"]), nat.ptr(p["absmax"]), None, None, None, nat.ptr(out), None, p["M"], p["N"],
       p["K"], p["bs"], nat.QT_ID[p["qt"]], nat.stream())
    return out


if __name__ == "__main__":
    main()
