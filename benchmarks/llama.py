"""bench.py --workload llama8b: BASELINE.json configs[4] -- a replica of the reference's
benchmarking/inference_benchmark.py (optimum-benchmark is not installed; there is no network):
a random-init Llama-3-8B (hidden 4096, intermediate 14336, 32 layers, 32 heads, 8 KV heads,
vocab 128256) in bf16 whose nn.Linear layers (all but lm_head) are swapped for Linear4bit
(NF4, compute bf16), batch 1, prompt 2048 random token ids; prefill time and decode tokens/s
(decode tokens / decode time, as optimum-benchmark reports it).
"""
from __future__ import annotations

import json
import time


def swap_linears(model, quant_type="nf4", compress_statistics=False, skip=("lm_head",)):
    import torch

    from bitsandbytes_b200.nn import Linear4bit, Params4bit

    n = 0
    for name, mod in list(model.named_modules()):
        for child_name, child in list(mod.named_children()):
            full = f"{name}.{child_name}" if name else child_name
            if isinstance(child, torch.nn.Linear) and not any(s in full for s in skip):
                q = Linear4bit(child.in_features, child.out_features, bias=child.bias is not None,
                               compute_dtype=torch.bfloat16, compress_statistics=compress_statistics,
                               quant_type=quant_type, device="meta")
                q.weight = Params4bit(child.weight.data, requires_grad=False, compress_statistics=compress_statistics,
                                      quant_type=quant_type, module=q)
                if child.bias is not None:
                    q.bias = child.bias
                q.weight = q.weight.to(child.weight.device)  # quantises on the GPU
                setattr(mod, child_name, q)
                child.weight = None
                n += 1
    return n


def run_llama8b(args, rank: int, world: int, local_rank: int) -> None:
    import torch

    if rank != 0:
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    print(json.dumps(measure_llama8b(dev, int(getattr(args, "layers", 32) or 32), args.steps)), flush=True)


def measure_llama8b(dev, layers: int = 32, steps: int = 64):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM

    class _A:
        pass

    args = _A()
    args.layers, args.steps = layers, steps
    # cuDNN's SDPA backend rebuilds its plan on the host for every new KV length (~12 ms per call in
    # decode); the flash / memory-efficient backends do not.  Attention is not part of the measured path.
    torch.backends.cuda.enable_cudnn_sdp(False)
    layers = int(getattr(args, "layers", 32) or 32)
    cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers, num_attention_heads=32,
                      num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, rms_norm_eps=1e-5,
                      rope_theta=500000.0, tie_word_embeddings=False)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    with torch.device(dev):
        model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    model.eval()
    n_swapped = swap_linears(model)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0

    prompt_len, new_tokens = 2048, max(64, min(args.steps, 256))
    ids = torch.randint(0, cfg.vocab_size, (1, prompt_len), device=dev)

    @torch.no_grad()
    def run_once(n_new):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = model(input_ids=ids, use_cache=True)
        past = out.past_key_values
        tok = out.logits[:, -1:].argmax(-1)
        e[1].record()
        for _ in range(n_new):
            out = model(input_ids=tok, past_key_values=past, use_cache=True)
            past = out.past_key_values
            tok = out.logits[:, -1:].argmax(-1)
        e[2].record()
        torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])

    run_once(4)  # warm-up (workspace allocation, autotuned attention kernels)
    prefill_ms, decode_ms = run_once(new_tokens)
    mode = "eager HF forward"

    # CUDA-graph decode: one captured decode step over a static KV cache, replayed per token
    # ("CUDA streams and graphs instead of a tracing compiler").  Falls back to the eager numbers
    # above if this transformers version cannot be captured.
    graph_info = None
    try:
        from transformers import StaticCache

        cache = StaticCache(config=cfg, max_cache_len=prompt_len + new_tokens + 16)
        tok_buf = torch.zeros((1, 1), dtype=torch.long, device=dev)
        pos_buf = torch.zeros((1,), dtype=torch.long, device=dev)
        side = torch.cuda.Stream()
        with torch.no_grad():
            out = model(input_ids=ids, past_key_values=cache, cache_position=torch.arange(prompt_len, device=dev),
                        use_cache=True)
            tok_buf.copy_(out.logits[:, -1:].argmax(-1))
            pos_buf.fill_(prompt_len)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):  # warm-up on the capture stream (library workspaces are per stream)
                    o = model(input_ids=tok_buf, past_key_values=cache, cache_position=pos_buf, use_cache=True)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                o = model(input_ids=tok_buf, past_key_values=cache, cache_position=pos_buf, use_cache=True)
                nxt = o.logits[:, -1:].argmax(-1)
                tok_buf.copy_(nxt)
                pos_buf.add_(1)
            torch.cuda.synchronize()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            pos_buf.fill_(prompt_len)
            e0.record()
            for _ in range(new_tokens):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            graph_ms = e0.elapsed_time(e1)
        graph_info = {"decode_ms_per_token": graph_ms / new_tokens, "tokens_per_s": new_tokens / (graph_ms * 1e-3)}
        eager_tps = new_tokens / (decode_ms * 1e-3)
        if graph_info["tokens_per_s"] > eager_tps:
            graph_info["eager_tokens_per_s"] = eager_tps
            decode_ms = graph_ms
            mode = "CUDA-graph decode step over a static KV cache"
    except Exception as exc:  # noqa: BLE001
        graph_info = {"error": repr(exc)[:300]}
    line = {
        "metric": "llama3_8b_nf4_decode_tokens_per_s", "value": new_tokens / (decode_ms * 1e-3), "unit": "tokens/s",
        "n_gpus": 1, "steps": new_tokens, "warmup": 4, "ms_per_step": decode_ms / new_tokens, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic (random-init weights, random ids)",
        "config": {"workload": "llama8b", "layers": layers, "linear4bit_layers": n_swapped, "quant_type": "nf4",
                   "batch": 1, "prompt_tokens": prompt_len, "new_tokens": new_tokens, "mode": mode},
        "cuda_graph": graph_info,
        "prefill_ms": prefill_ms, "prefill_tokens_per_s": prompt_len / (prefill_ms * 1e-3), "build_s": t_build,
        "published_reference": "H100 SXM, bitsandbytes 0.45: Llama 3.1 8B NF4 bs=1 30.14 tok/s (benchmarking/README.md:91)",
    }
    del model
    torch.cuda.empty_cache()
    return line
