"""Dev tool (GPU, not a pytest file): where one Orpheus-3B batch-32 decode step spends its time, IN SITU.

Captures CUDA graphs of one decode step with pieces removed (results are garbage, timing is not) and replays each
back to back: full step, LM forward only, no attention, no projections, ... The difference to the full step is
the marginal cost of the piece inside the real pipeline (PDL overlap included), which is what ncu's serialised
per-launch times cannot show.

    python tests/ablate_step.py [kv_len] [batch]
Prints one JSON line per variant.
"""
import json
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402
from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights  # noqa: E402
from vox_serve_b200.model.orpheus import synthetic_state_dict  # noqa: E402

BF = torch.bfloat16


def main():
    kv_len = int(sys.argv[1]) if len(sys.argv) > 1 else 728
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = "cuda"
    d = LlamaDims.orpheus_3b()
    w = LlamaWeights.from_state_dict(synthetic_state_dict(d, 0, dev), d, dev)
    ps = 128
    pages_req = (kv_len + ps) // ps + 1
    n_pages = B * pages_req
    kv = (torch.randn(d.num_hidden_layers, n_pages, 2, ps, d.num_key_value_heads, d.head_dim, device=dev) * 0.5).to(BF)
    eng = LlamaEngine(w, kv, ps, max_rows=64)
    npg = (kv_len + ps - 1) // ps
    indptr = torch.arange(B + 1, dtype=torch.int32, device=dev) * npg
    perm = torch.randperm(n_pages, device=dev).to(torch.int32)
    indices = torch.cat([perm[r * pages_req: r * pages_req + npg] for r in range(B)]).contiguous()
    last = torch.full((B,), kv_len - (npg - 1) * ps, dtype=torch.int32, device=dev)
    ops.plan_rows(eng.plan, None, indptr, indices, last, B, B, ps, eng.chunk)
    ids = torch.randint(128266, 156000, (B,), dtype=torch.int32, device=dev)
    pos = torch.full((B,), kv_len - 1, dtype=torch.int32, device=dev)
    rep = torch.zeros(B, 1, 1, d.vocab_size, dtype=torch.uint8, device=dev)
    rng = torch.tensor([1, 0, 0], dtype=torch.int64, device=dev)
    out_ids = torch.zeros(B, dtype=torch.int64, device=dev)
    hq, hkv, D, H, I = d.num_attention_heads, d.num_key_value_heads, d.head_dim, d.hidden_size, d.intermediate_size
    tiles_h = (H + 127) // 128

    def sample():
        ops.sample(eng.logits[:B], "top_p", rep_cache=rep, penalty=1.3, top_p=0.8, temperature=0.6, rng_state=rng,
                   out=out_ids)
        ops.update_repetition_cache(rep, out_ids.view(B, 1), -1)

    def layers(skip=()):
        hidden, q, attn, act = eng.hidden[:B], eng.q[:B], eng.attn[:B], eng.act[:B]
        ssq = eng.ssq[: tiles_h * B].view(tiles_h, B)
        cs = ops.rope_table(pos, eng.freq, D, out=eng.rope_cs[:B])
        ops.row_ssq(hidden, out=ssq[0])
        for i, L in enumerate(w.layers):
            if "qkv" not in skip:
                ops.proj_norm_qkv_rope_append(hidden, ssq, tiles_h, L["ln1"], d.rms_norm_eps, L["qkv"], kv[i], cs, eng.plan,
                                              hq, hkv, D, eng.fsplit_qkv, q_out=q)
            if "attn" not in skip:
                ops.paged_attn(q, kv, i * n_pages, eng.plan, B, hkv, ps, eng.chunk, eng.attn_ws, out=attn,
                               grid_ctas=eng.attn_grid)
            if "o" not in skip:
                ops.proj_residual(attn.view(B, hq * D), L["o"], hidden, eng.fsplit_o, hidden_out=hidden,
                                  ssq_out=ssq)
            if "gu" not in skip:
                ops.proj_norm_gateup_silu(hidden, ssq, tiles_h, L["ln2"], d.rms_norm_eps, L["gu"], eng.gu_half, I, out=act)
            if "down" not in skip:
                ops.proj_residual(act, L["down"], hidden, eng.fsplit_down, hidden_out=hidden, ssq_out=ssq)

    def step(skip=()):
        ops.embedding(w.embed, ids, out=eng.hidden[:B])
        layers(skip)
        ops.rmsnorm(eng.hidden[:B], w.norm, d.rms_norm_eps, out=eng.normed[:B])
        if "lm_head" not in skip:
            ops.gemm(eng.normed[:B], w.lm_head, mode=0, out=eng.logits[:B])
        if "sample" not in skip:
            sample()

    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    variants = {
        "full": lambda: step(),
        "no_sample": lambda: step(("sample",)),
        "no_lm_head_sample": lambda: step(("sample", "lm_head")),
        "no_attn": lambda: step(("attn",)),
        "no_qkv": lambda: step(("qkv",)),
        "no_o": lambda: step(("o",)),
        "no_gu": lambda: step(("gu",)),
        "no_down": lambda: step(("down",)),
        "attn_only": lambda: step(("qkv", "o", "gu", "down", "lm_head", "sample")),
        "proj_only": lambda: step(("attn", "sample")),
        "unfused_forward": None,
    }

    def unfused():
        eng.forward(ids, pos, B)

    def fused():
        eng.force_unfused = False
        eng.forward(ids, pos, B)
        eng.force_unfused = True

    def unfused_rows():
        eng.tiled_acts = False
        eng.forward(ids, pos, B)
        eng.tiled_acts = True

    variants["unfused_rows_forward"] = unfused_rows
    variants["unfused_forward"] = unfused
    variants["fused_forward"] = fused

    res = {}
    if only:
        variants = {k: v for k, v in variants.items() if k in only}
    for name, fn in variants.items():
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()                                   # warm-up (also sets function attributes outside capture)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps
        print(json.dumps({"variant": name, "ms": round(res[name], 4), "kv_len": kv_len, "batch": B}), flush=True)
    if "full" not in res:
        return
    full = res["full"]
    w_bytes = w.streamed_bytes_per_step()
    kv_bytes = d.num_hidden_layers * B * kv_len * 2 * hkv * D * 2
    print(json.dumps({"summary": {k: round(full - v, 4) for k, v in res.items() if k.startswith("no_")},
                      "full_ms": round(full, 4), "weights_GB": w_bytes / 1e9, "kv_GB": kv_bytes / 1e9,
                      "step_GBps": (w_bytes + kv_bytes) / full / 1e6}))


if __name__ == "__main__":
    main()
