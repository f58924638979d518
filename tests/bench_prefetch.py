"""Dev tool (GPU): does pulling the next projection's weights into L2 from inside a GEMM shorten the next GEMM?"""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402

BF = torch.bfloat16
dev = "cuda"
T = 32
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
xa = torch.randn(T, 3072, device=dev).to(BF)
wa = (torch.randn(16384, 3072, device=dev) * 0.02).to(BF)      # gate/up: 100 MB
oa = torch.empty(T, 8192, dtype=BF, device=dev)
for name, N, K, split in (("down", 3072, 8192, 6), ("o", 3072, 3072, 6), ("qkv", 5120, 3072, 3)):
    xb = torch.randn(T, K, device=dev).to(BF)
    wb = (torch.randn(N, K, device=dev) * 0.02).to(BF)
    ob = torch.empty(split, T, N, dtype=torch.float32, device=dev)
    for pf in (False, True):
        ts_a, ts_b = [], []
        for it in range(12):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            ops.gemm(xa, wa, mode=2, out=oa, prefetch=wb if pf else None)
            e1.record()
            ops.gemm(xb, wb, mode=1, split_k=split, out=ob)
            e2.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts_a.append(e0.elapsed_time(e1) * 1e3)
                ts_b.append(e1.elapsed_time(e2) * 1e3)
        ts_a.sort(); ts_b.sort()
        print(f"{name:5s} prefetch={pf!s:5s}  A(gate_up) {ts_a[len(ts_a)//2]:7.2f} us   B({name}) {ts_b[len(ts_b)//2]:7.2f} us   "
              f"sum {ts_a[len(ts_a)//2] + ts_b[len(ts_b)//2]:7.2f}", flush=True)
