import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import LlamaConfig, LlamaForCausalLM
from benchmarks.llama import swap_linears
dev = torch.device("cuda")
cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=4, num_attention_heads=32,
                  num_key_value_heads=8, vocab_size=128256, max_position_embeddings=8192, tie_word_embeddings=False)
with torch.device(dev):
    model = LlamaForCausalLM(cfg).to(torch.bfloat16)
model.eval(); swap_linears(model)
ids = torch.randint(0, cfg.vocab_size, (1, 2048), device=dev)
with torch.no_grad():
    out = model(input_ids=ids, use_cache=True); past = out.past_key_values; tok = out.logits[:, -1:].argmax(-1)
    for _ in range(3):
        out = model(input_ids=tok, past_key_values=past, use_cache=True); past = out.past_key_values
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        out = model(input_ids=tok, past_key_values=past, use_cache=True); past = out.past_key_values
    torch.cuda.synchronize(); print("ms/step (4 layers):", (time.perf_counter() - t0) * 100)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            out = model(input_ids=tok, past_key_values=past, use_cache=True); past = out.past_key_values
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=18, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
    # raw timing of a single linear call
    lin = model.model.layers[0].self_attn.q_proj
    x = torch.randn(1, 1, 4096, device=dev, dtype=torch.bfloat16)
    for _ in range(5): lin(x)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): lin(x)
    torch.cuda.synchronize(); print("us per Linear4bit call:", (time.perf_counter() - t0) / 200 * 1e6)
