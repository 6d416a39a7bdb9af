"""Kernel-level breakdown of one CUDA-graph decode step of the llama8b workload (torch profiler / CUPTI).
usage: profile_llama_decode.py [layers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from transformers import LlamaConfig, LlamaForCausalLM, StaticCache  # noqa: E402

from benchmarks.llama import swap_linears  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
torch.backends.cuda.enable_cudnn_sdp(False)
cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers, num_attention_heads=32,
                  num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, rms_norm_eps=1e-5,
                  rope_theta=500000.0, tie_word_embeddings=False)
torch.manual_seed(0)
with torch.device(dev):
    model = LlamaForCausalLM(cfg).to(torch.bfloat16)
model.eval()
swap_linears(model)
prompt_len, new_tokens = 2048, 16
ids = torch.randint(0, cfg.vocab_size, (1, prompt_len), device=dev)
cache = StaticCache(config=cfg, max_cache_len=prompt_len + 64 + 16)
tok_buf = torch.zeros((1, 1), dtype=torch.long, device=dev)
pos_buf = torch.zeros((1,), dtype=torch.long, device=dev)
side = torch.cuda.Stream()
with torch.no_grad():
    out = model(input_ids=ids, past_key_values=cache, cache_position=torch.arange(prompt_len, device=dev), use_cache=True)
    tok_buf.copy_(out.logits[:, -1:].argmax(-1))
    pos_buf.fill_(prompt_len)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
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
    e0.record()
    for _ in range(new_tokens):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"layers={layers}: graph decode {e0.elapsed_time(e1) / new_tokens * 1e3:.0f} us per token", flush=True)
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(4):
            g.replay()
        torch.cuda.synchronize()
    rows = []
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0)
        if t > 0 and ev.device_type.name != "CPU":
            rows.append((t / 4.0, ev.count / 4.0, ev.key))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"kernel time per token {tot:.0f} us over {sum(r[1] for r in rows):.0f} launches")
    for t, c, k in rows[:28]:
        print(f"{t:9.1f} us  x{c:6.1f}  {t / max(c, 1):7.2f} us each  {k[:110]}")
