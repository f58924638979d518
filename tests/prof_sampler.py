"""Dev tool (GPU): a few sampler launches at the bench's shape (32 rows x 156940 logits, top-p 0.8, T 0.6, penalty 1.3), for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402

B, V, dev = 32, 156940, "cuda"
logits = (torch.randn(B, V, device=dev) * 1.1).to(torch.bfloat16)
rep = (torch.rand(B, 1, 1, V, device=dev) < 0.002).to(torch.uint8)
rng = torch.tensor([1, 0, 0], dtype=torch.int64, device=dev)
out = torch.zeros(B, dtype=torch.int64, device=dev)
rows = torch.arange(B, dtype=torch.int32, device=dev)
for _ in range(4):
    ops.sample(logits, "top_p", rep_cache=rep, penalty=1.3, top_p=0.8, temperature=0.6, rng_state=rng, out=out,
               cache_rows=rows)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.sample(logits, "top_p", rep_cache=rep, penalty=1.3, top_p=0.8, temperature=0.6, rng_state=rng, out=out,
               cache_rows=rows)
e1.record()
torch.cuda.synchronize()
print("sampler us/launch:", e0.elapsed_time(e1) * 1e3 / 20, "ids", out[:6].tolist())
