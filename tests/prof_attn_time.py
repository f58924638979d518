"""Timing probe (GPU): the decode attention kernel as bench.py's roofline_attention measures it -- the 28 launches of
one decode step (one per layer, each over its own layer's KV so nothing is L2-resident) captured in a CUDA graph,
replayed back to back, CUDA-event timed.  One JSON line per kv length.

    python tests/prof_attn_time.py 200 500 900
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402

PEAK = 6553.3


def main():
    kvs = [int(a) for a in sys.argv[1:]] or [200, 500, 900]
    B, hq, hkv, D, ps, L = 32, 24, 8, 128, 128, 28
    dev = "cuda"
    for kvlen in kvs:
        # the bench's steady-state mix around this mean: groups of four requests staggered by kv/4 steps
        lens = [max(1, kvlen + (i // 4 - 4) * (kvlen // 10)) for i in range(B)]
        pages_req = [(n + ps - 1) // ps for n in lens]
        n_pages = sum(pages_req) + 8
        cache = torch.randn(L, n_pages, 2, ps, hkv, D, device=dev).to(torch.bfloat16)
        perm = torch.randperm(n_pages, device=dev).to(torch.int32)
        indptr = torch.tensor([0] + list(torch.tensor(pages_req).cumsum(0)), dtype=torch.int32, device=dev)
        indices = perm[: sum(pages_req)].contiguous()
        last = torch.tensor([n - (p - 1) * ps for n, p in zip(lens, pages_req)], dtype=torch.int32, device=dev)
        plan = ops.RowPlan(B, dev)
        TOK = ops.attn_chunk_tokens(ps, hkv)
        ops.plan_rows(plan, None, indptr, indices, last, B, B, ps, TOK)
        ws = ops.AttnWorkspace(B, hq, hkv, D, dev)
        q = torch.randn(B, hq, D, device=dev).to(torch.bfloat16)
        o = torch.empty_like(q)

        def step():
            for layer in range(L):
                ops.paged_attn(q, cache, layer * n_pages, plan, B, hkv, ps, TOK, ws, out=o)
        step()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                step()
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (20 * L)
        nbytes = sum(lens) * hkv * D * 2 * 2 + 2 * B * hq * D * 2
        print(json.dumps({"mean_kv": sum(lens) / B, "us_per_launch": round(us, 2), "bytes": nbytes,
                          "gb_s": round(nbytes / us / 1e3, 1), "frac": round(nbytes / us / 1e3 / PEAK, 3)}), flush=True)
        del cache


if __name__ == "__main__":
    main()
