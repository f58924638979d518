"""Dev tool (GPU): run the medium Orpheus-shaped model's first prefill step op by op on the device and on the
oracle, and print the first op whose output leaves bf16 agreement.  python tests/debug_layers.py [T]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lm_ops, orpheus as oorph  # noqa: E402

BF = torch.bfloat16


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-9)).item()


def main():
    from vox_serve_b200 import ops
    from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights

    T = int(sys.argv[1]) if len(sys.argv) > 1 else 133
    page = 128
    dims = oorph.OrpheusDims.tiny(hidden_size=1536, num_hidden_layers=3, num_attention_heads=12,
                                  num_key_value_heads=4, intermediate_size=2048, vocab_size=10 + 7 * 4096)
    w = oorph.synth_weights(dims, seed=4)
    ld = LlamaDims(dims.hidden_size, dims.num_hidden_layers, dims.num_attention_heads, dims.num_key_value_heads,
                   dims.head_dim, dims.intermediate_size, dims.vocab_size, dims.rms_norm_eps, dims.rope_theta,
                   dims.rope_factor, dims.low_freq_factor, dims.high_freq_factor, dims.old_context_len)
    gw = LlamaWeights.from_state_dict(w, ld)
    n_pages = 16
    kv = torch.zeros(dims.num_hidden_layers, n_pages, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF,
                     device="cuda")
    eng = LlamaEngine(gw, kv, page, max_rows=256)
    g = torch.Generator().manual_seed(21)
    ids = torch.randint(10, dims.vocab_size, (T,), generator=g)
    pos = torch.arange(T, dtype=torch.int32)
    npg = (T + page - 1) // page
    qo, ip, idx, last = [0, T], [0, npg], list(range(npg)), [T % page or page]
    i32 = lambda x: torch.tensor(x, dtype=torch.int32, device="cuda")
    ops.plan_rows(eng.plan, i32(qo), i32(ip), i32(idx), i32(last), 1, T, page, eng.chunk)
    # ---- oracle, op by op ----
    okv = torch.zeros(dims.num_hidden_layers, n_pages, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF)
    wr = lm_ops.PagedWrapperCPU("prefill", page)
    wr.plan(qo, ip, idx, last)
    d, hq, hkv, D, H, I = ld, ld.num_attention_heads, ld.num_key_value_heads, ld.head_dim, ld.hidden_size, ld.intermediate_size
    R = T
    d_ids, d_pos = ids.to(torch.int32).cuda(), pos.cuda()
    hidden, normed = eng.hidden[:R], eng.normed[:R]
    ops.embedding(gw.embed, d_ids, out=hidden)
    ops.rmsnorm(hidden, gw.layers[0]["ln1"], d.rms_norm_eps, out=normed)
    h = F.embedding(ids, w["model.embed_tokens.weight"])
    print("embed", rel(hidden, h))
    s_qkv, s_o, s_dn = eng._split(eng.split_qkv, R), eng._split(eng.split_o, R), eng._split(eng.split_down, R)
    print("splits", s_qkv, s_o, s_dn, "t_tile", ops.gemm_t_tile(R))
    q, attn, act = eng.q[:R], eng.attn[:R], eng.act[:R]
    for i, L in enumerate(gw.layers):
        n = oorph.layer_names(i)
        x = lm_ops.rms_norm(h, w[n["ln1"]], dims.rms_norm_eps)
        print(i, "ln1", rel(normed, x))
        p = ops.gemm(normed, L["qkv"], mode=1, split_k=s_qkv, out=eng._partials(s_qkv, R, eng.qkv_w))
        oq = F.linear(x, w[n["q"]]); ok = F.linear(x, w[n["k"]]); ov = F.linear(x, w[n["v"]])
        ref_qkv = torch.cat((oq, ok, ov), -1)
        print(i, "qkv gemm", rel(p.sum(0), ref_qkv))
        ops.qkv_rope_append(p, kv[i], d_pos, eng.freq, eng.plan, hq, hkv, D, q_out=q)
        rq, rk = lm_ops.apply_rope_pos_ids(oq.view(R, hq, D), ok.view(R, hkv, D), pos, rope_scale=dims.rope_factor,
                                           rope_theta=dims.rope_theta, low_freq_factor=dims.low_freq_factor,
                                           high_freq_factor=dims.high_freq_factor, old_context_len=dims.old_context_len)
        print(i, "rope q", rel(q, rq))
        wr.set_kv_cache(okv[i], rk, ov.view(R, hkv, D))
        print(i, "kv cache", rel(kv[i], okv[i]))
        ops.paged_attn(q, eng.kv_map, i * eng.pages_per_layer, eng.plan, R, hkv, page, eng.chunk, eng.attn_ws,
                       out=attn, grid_ctas=eng.attn_grid)
        a = wr.run(rq, okv[i]).reshape(R, -1)
        e = (attn.view(R, -1).float().cpu() - a.float()).abs().amax(dim=1)
        print(i, "attn", rel(attn.view(R, -1), a), "worst rows", torch.topk(e, 5).indices.tolist())
        p = ops.gemm(attn.view(R, hq * D), L["o"], mode=1, split_k=s_o, out=eng._partials(s_o, R, H))
        o = F.linear(a, w[n["o"]])
        print(i, "o gemm", rel(p.sum(0), o))
        ops.reduce_residual_rmsnorm(p, hidden, L["ln2"], d.rms_norm_eps, hidden_out=hidden, normed_out=normed)
        h = h + o
        x = lm_ops.rms_norm(h, w[n["ln2"]], dims.rms_norm_eps)
        print(i, "hidden", rel(hidden, h), "ln2", rel(normed, x))
        ops.gemm(normed, L["gu"], mode=2, out=act)
        gg = F.silu(F.linear(x, w[n["gate"]])) * F.linear(x, w[n["up"]])
        e = (act.float().cpu() - gg.float()).abs()
        print(i, "gate_up", rel(act, gg), "worst rows", torch.topk(e.amax(1), 5).indices.tolist(),
              "worst cols", torch.topk(e.amax(0), 5).indices.tolist())
        p = ops.gemm(act, L["down"], mode=1, split_k=s_dn, out=eng._partials(s_dn, R, H))
        dn = F.linear(gg, w[n["down"]])
        print(i, "down gemm", rel(p.sum(0), dn))
        nxt = gw.layers[i + 1]["ln1"] if i + 1 < len(gw.layers) else gw.norm
        ops.reduce_residual_rmsnorm(p, hidden, nxt, d.rms_norm_eps, hidden_out=hidden, normed_out=normed)
        h = h + dn
        print(i, "hidden2", rel(hidden, h))
    x = lm_ops.rms_norm(h, w["model.norm.weight"], dims.rms_norm_eps)
    print("final norm", rel(normed, x))
    last_rows = i32([T - 1])
    xl = ops.gather_rows(normed, last_rows, out=eng.last_normed[:1])
    print("gather", rel(xl, x[T - 1:T]))
    lg = ops.gemm(xl, gw.lm_head, mode=0, out=eng.logits[:1])
    ref = F.linear(x[T - 1:T], w["lm_head.weight"])
    print("logits", rel(lg, ref), "argmax", lg.float().argmax().item(), ref.float().argmax().item())
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
