"""A few Linear8bitLt forwards at the C3 shape (4096 x 11008, 4096 tokens, threshold 6, 5 outlier columns):
for the ncu launch list of the LLM.int8() forward.  usage: run_i8_fwd.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bitsandbytes_b200 as bnb  # noqa: E402

K, N, M = 4096, 11008, 4096
dev = torch.device("cuda", 0)
torch.manual_seed(0)
lin = torch.nn.Linear(K, N, bias=False)
layer = bnb.nn.Linear8bitLt(K, N, bias=False, has_fp16_weights=False, threshold=6.0)
layer.load_state_dict(lin.state_dict())
layer = layer.to(dev).eval()
cols = torch.randint(0, K, (5,), generator=torch.Generator().manual_seed(1)).tolist()
x = torch.randn(M, K, device=dev, dtype=torch.float16)
x[:, cols] = 8.0
with torch.no_grad():
    for _ in range(4):
        y = layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = layer(x)
    e1.record()
    torch.cuda.synchronize()
print(f"Linear8bitLt forward {e0.elapsed_time(e1) * 100:.1f} us per call (eager, 5 outlier columns)")
