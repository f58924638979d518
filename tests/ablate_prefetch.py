"""Dev tool (GPU): one Orpheus-3B decode forward (32 rows) as a CUDA graph under different L2-prefetcher settings.
python tests/ablate_prefetch.py [kv_len]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402
from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights  # noqa: E402
from vox_serve_b200.model.orpheus import synthetic_state_dict  # noqa: E402


def main():
    kv_len = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    B, dev, ps = 32, "cuda", 128
    d = LlamaDims.orpheus_3b()
    w = LlamaWeights.from_state_dict(synthetic_state_dict(d, 0, dev), d, dev)
    pages_req = (kv_len + ps) // ps + 1
    n_pages = B * pages_req
    kv = (torch.randn(d.num_hidden_layers, n_pages, 2, ps, d.num_key_value_heads, d.head_dim, device=dev) * 0.5).to(torch.bfloat16)
    eng = LlamaEngine(w, kv, ps, max_rows=64)
    npg = (kv_len + ps - 1) // ps
    indptr = torch.arange(B + 1, dtype=torch.int32, device=dev) * npg
    perm = torch.randperm(n_pages, device=dev).to(torch.int32)
    indices = torch.cat([perm[r * pages_req: r * pages_req + npg] for r in range(B)]).contiguous()
    last = torch.full((B,), kv_len - (npg - 1) * ps, dtype=torch.int32, device=dev)
    ops.plan_rows(eng.plan, None, indptr, indices, last, B, B, ps, eng.chunk)
    ids = torch.randint(128266, 156000, (B,), dtype=torch.int32, device=dev)
    pos = torch.full((B,), kv_len - 1, dtype=torch.int32, device=dev)
    configs = [
        ("off", dict()),
        ("inline: gu 64 + next qkv/o", dict(pf_inline=True)),
        ("inline: gu 32 + next qkv/o", dict(pf_inline=True, pf_inline_gu_mb=32)),
        ("inline: gu 96 + next qkv/o", dict(pf_inline=True, pf_inline_gu_mb=96)),
        ("inline: next qkv/o only", dict(pf_inline=True, pf_inline_gu_mb=0)),
        ("inline: gu 64 + qkv/o + down 48 at rope", dict(pf_inline=True, pf_inline_down_mb=48)),
        ("inline: gu 48 + qkv/o + down 32 at rope", dict(pf_inline=True, pf_inline_gu_mb=48, pf_inline_down_mb=32)),
        ("polling kernel, 8 CTAs, weights", dict(l2_prefetch=True, l2_prefetch_kv=False, l2_prefetch_ctas=8)),
        ("polling kernel, 8 CTAs, weights + KV", dict(l2_prefetch=True, l2_prefetch_kv=True, l2_prefetch_ctas=8)),
        ("off (again)", dict()),
    ]
    defaults = dict(l2_prefetch=False, l2_prefetch_kv=True, l2_window_mb=64, l2_prefetch_ctas=0, l2_prefetch_flags=0,
                    l2_prefetch_kernel=True, pf_inline=False, pf_inline_gu_mb=64, pf_inline_down_mb=0)
    for name, cfg in configs:
        for k, v in {**defaults, **cfg}.items():
            setattr(eng, k, v)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            eng.forward(ids, pos, B)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                eng.forward(ids, pos, B)
        torch.cuda.current_stream().wait_stream(s)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        first = e0.elapsed_time(e1)
        reps = 20 if first < 10 else 2
        t0 = time.time()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:34s} {e0.elapsed_time(e1) / reps:8.3f} ms / forward  (first replay {first:.3f} ms, {reps} reps, "
              f"{time.time() - t0:.2f} s wall)", flush=True)


if __name__ == "__main__":
    main()
