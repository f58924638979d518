"""Micro-benchmarks of the hot kernels at Orpheus-3B decode shapes (run on the GPU box; not a pytest file).
Prints one JSON line per kernel with achieved GB/s against the algorithmic bytes."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402

BF = torch.bfloat16


def timeit(fn, iters=20, warm=5, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3, ts[0] * 1e-3


def main():
    dev = "cuda"
    which = set(sys.argv[1:]) or {"gemm", "attn", "sampler"}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = []
    # ---- GEMMs, T = 32 ----
    T = 32
    for name, N, K, mode in () if "gemm" not in which else (("qkv", 5120, 3072, 1), ("o", 3072, 3072, 1), ("gate_up", 16384, 3072, 2),
                             ("down", 3072, 8192, 1), ("lm_head", 156940, 3072, 0)):
        x = torch.randn(T, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * 0.02).to(BF)
        for split in ([1] if mode != 1 else sorted({1, ops.choose_split_k(N, K, T), 2, 4, 6, 8})):
            if mode == 1:
                o = torch.empty(split, T, N, dtype=torch.float32, device=dev)
            else:
                o = torch.empty(T, N // 2 if mode == 2 else N, dtype=BF, device=dev)
            fn = lambda: ops.gemm(x, w, mode=mode, split_k=split, out=o)
            med, best = timeit(fn, flush=flush)
            byt = N * K * 2
            out.append(dict(kernel=f"gemm_{name}", split_k=split, ms=med * 1e3, best_ms=best * 1e3,
                            GBs=byt / med / 1e9, best_GBs=byt / best / 1e9))
            print(json.dumps(out[-1]), flush=True)
    # ---- paged decode attention ----
    hq, hkv, D, ps = 24, 8, 128, 128
    for kvlen in () if "attn" not in which else (160, 250, 480, 728, 890, 1333):
        B = 32
        n_pages_req = (kvlen + ps - 1) // ps
        n_pages = B * n_pages_req + 8
        cache = torch.randn(1, n_pages, 2, ps, hkv, D, device=dev).to(BF)
        indptr = torch.arange(B + 1, dtype=torch.int32, device=dev) * n_pages_req
        indices = torch.randperm(n_pages, device=dev)[: B * n_pages_req].to(torch.int32)
        last = torch.full((B,), kvlen - (n_pages_req - 1) * ps, dtype=torch.int32, device=dev)
        mc = B * ((kvlen + 63) // 64)
        plan = ops.RowPlan(B, dev, mc)
        TOK = ops.attn_chunk_tokens(ps, hkv)
        ops.plan_rows(plan, None, indptr, indices, last, B, B, ps, TOK)
        ws = ops.AttnWorkspace(B, hq, hkv, D, dev, grid_ctas=296)
        kv_map = ops.tensor_map_kv(cache, TOK)
        q = torch.randn(B, hq, D, device=dev).to(BF)
        o = torch.empty_like(q)
        byt = B * kvlen * 2 * hkv * D * 2 + 2 * q.numel() * 2
        for grid in (74, 148, 296):
            fn = lambda: ops.paged_attn(q, kv_map, 0, plan, B, hkv, ps, TOK, ws, out=o, grid_ctas=grid)
            med, best = timeit(fn, flush=flush)
            out.append(dict(kernel="paged_attn_decode", kv_len=kvlen, grid=grid, ms=med * 1e3, best_ms=best * 1e3,
                            GBs=byt / med / 1e9, best_GBs=byt / best / 1e9))
            print(json.dumps(out[-1]), flush=True)
    # ---- sampler ----
    V = 156940
    logits = (torch.randn(32, V, device=dev) * 4).to(BF)
    cache = torch.zeros(32, 1, 1, V, dtype=torch.bool, device=dev)
    for strat, kw in () if "sampler" not in which else (("greedy", {}), ("top_p", dict(top_p=0.8, temperature=0.6))):
        fn = lambda: ops.sample(logits, strat, rep_cache=cache, penalty=1.3, **kw)
        med, best = timeit(fn)
        out.append(dict(kernel=f"sample_{strat}", ms=med * 1e3, best_ms=best * 1e3))
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
