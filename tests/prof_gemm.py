"""Dev tool (GPU): a few launches of the gate/up projection at Orpheus decode geometry (32 rows, tiled activations,
56 + 56-row tiles -> 147 CTAs), for ncu."""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402

BF = torch.bfloat16
T, H, I, dev = 32, 3072, 8192, "cuda"
h = ops.gate_up_tile_half(I)
wg = (torch.randn(I, H, device=dev) * 0.02).to(BF)
wu = (torch.randn(I, H, device=dev) * 0.02).to(BF)
w = ops.pack_weight(ops.interleave_gate_up(wg, wu, h), 2 * h)
x = ops.rmsnorm((torch.randn(T, H, device=dev) * 2).to(BF), torch.ones(H, device=dev, dtype=BF), 1e-5,
                out=ops.TiledAct(T, H, dev))
act = ops.TiledAct(T, I, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(4):
    flush.zero_()
    ops.gemm(x, w, mode=2, out=act, tile_rows=2 * h, n_out=I)
torch.cuda.synchronize()
