"""Dev tool (GPU): a few paged-attention launches at Orpheus decode geometry, for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402

kvlen = int(sys.argv[1]) if len(sys.argv) > 1 else 728
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 148
B, hq, hkv, D, ps = 32, 24, 8, 128, 128
dev = "cuda"
n_pages_req = (kvlen + ps - 1) // ps
n_pages = B * n_pages_req + 8
cache = torch.randn(1, n_pages, 2, ps, hkv, D, device=dev).to(torch.bfloat16)
indptr = torch.arange(B + 1, dtype=torch.int32, device=dev) * n_pages_req
indices = torch.randperm(n_pages, device=dev)[: B * n_pages_req].to(torch.int32)
last = torch.full((B,), kvlen - (n_pages_req - 1) * ps, dtype=torch.int32, device=dev)
plan = ops.RowPlan(B, dev)
TOK = ops.attn_chunk_tokens(ps, hkv)
ops.plan_rows(plan, None, indptr, indices, last, B, B, ps, TOK)
ws = ops.AttnWorkspace(B, hq, hkv, D, dev, grid_ctas=grid)
kv_map = ops.tensor_map_kv(cache, TOK)
q = torch.randn(B, hq, D, device=dev).to(torch.bfloat16)
o = torch.empty_like(q)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    ops.paged_attn(q, kv_map, 0, plan, B, hkv, ps, TOK, ws, out=o, grid_ctas=grid)
torch.cuda.synchronize()
