"""Dev tool (GPU): in-situ timeline of one CUDA-graph replay of the decode step (block 0 of every instrumented kernel
logs %globaltimer through vb_set_trace).  python tests/trace_step.py [kv_len] [mode: unfused|fused|attn|gu] [first] [count]"""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import _lib, ops  # noqa: E402
from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights  # noqa: E402
from vox_serve_b200.model.orpheus import synthetic_state_dict  # noqa: E402

NAMES = {5: "reduce+norm", 6: "rope+append", 24: " rope:dep-released", 22: " attn:first-tile", 23: " attn:last-tile", 1: "gemm", 2: "attn", 4: "sample", 10: "rope_tab", 20: " gemm:dep-released", 21: " gemm:acc-done"}


def segments(recs, n_layers):
    """Critical-path segments of the default decode layer, averaged over layers 2 .. n-2 (us)."""
    def ts(kid, end=False):
        return sorted((r[1] if end else r[0]) for r in recs if r[2] == kid)

    rel, acc, gend = ts(20), ts(21), ts(1, True)
    rrel, rend, af, al, aend, nend = ts(24), ts(6, True), ts(22), ts(23), ts(2, True), ts(5, True)
    if not (len(rel) >= 4 * n_layers + 1 and len(nend) >= 2 * n_layers and len(af) >= n_layers):
        print("segments: unexpected record counts", len(rel), len(acc), len(nend), len(af), len(rrel))
        return
    names = ["qkv rel->acc", "acc->rope rel", "rope rel->attn first tile", "attn first->last tile", "attn last tile->end",
             "attn end->O rel", "O rel->acc", "O acc->reduce end", "reduce end->GU rel", "GU rel->acc", "GU acc->down rel",
             "down rel->acc", "down acc->reduce end", "reduce end->next qkv rel"]
    tot = [0.0] * len(names)
    layers = range(2, n_layers - 2)
    for L in layers:
        q = 4 * L
        pts = [rel[q], acc[q], rrel[L], af[L], al[L], aend[L], rel[q + 1], acc[q + 1], nend[2 * L], rel[q + 2], acc[q + 2],
               rel[q + 3], acc[q + 3], nend[2 * L + 1], rel[q + 4]]
        for i in range(len(names)):
            tot[i] += (pts[i + 1] - pts[i]) / 1e3
    n = len(layers)
    print("per-layer critical path (us, mean of layers 2..%d): total %.1f" % (n_layers - 3, sum(tot) / n))
    for nm, v in zip(names, tot):
        print(f"   {nm:28s} {v / n:6.2f}")


def main():
    kv_len = int(sys.argv[1]) if len(sys.argv) > 1 else 728
    mode = sys.argv[2] if len(sys.argv) > 2 else "unfused"
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    count = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    B, dev, ps = 32, "cuda", 128
    d = LlamaDims.orpheus_3b()
    w = LlamaWeights.from_state_dict(synthetic_state_dict(d, 0, dev), d, dev)
    pages_req = (kv_len + ps) // ps + 1
    n_pages = B * pages_req
    kv = (torch.randn(d.num_hidden_layers, n_pages, 2, ps, d.num_key_value_heads, d.head_dim, device=dev) * 0.5).to(torch.bfloat16)
    eng = LlamaEngine(w, kv, ps, max_rows=64)
    eng.force_unfused = mode == "unfused"
    npg = (kv_len + ps - 1) // ps
    indptr = torch.arange(B + 1, dtype=torch.int32, device=dev) * npg
    perm = torch.randperm(n_pages, device=dev).to(torch.int32)
    indices = torch.cat([perm[r * pages_req: r * pages_req + npg] for r in range(B)]).contiguous()
    last = torch.full((B,), kv_len - (npg - 1) * ps, dtype=torch.int32, device=dev)
    ops.plan_rows(eng.plan, None, indptr, indices, last, B, B, ps, eng.chunk)
    ids = torch.randint(128266, 156000, (B,), dtype=torch.int32, device=dev)
    pos = torch.full((B,), kv_len - 1, dtype=torch.int32, device=dev)
    cap = 4096
    buf = torch.zeros(4 + 4 * cap + 10 * 256, dtype=torch.int64, device=dev)
    buf[1] = cap
    buf[2] = 4 + 4 * cap
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    def run():
        if mode == "attn":      # the 28 attention launches of one step, back to back (what bench.py's roofline_attention times)
            for i in range(d.num_hidden_layers):
                eng.attention_only(i, B, eng.plan)
        elif mode == "gu":      # only the gate/up projections, back to back (instruction cache stays warm)
            for L in w.layers:
                ops.gemm(eng.normed_t.view_rows(B), L["gu"], mode=2, out=eng.act_t.view_rows(B), tile_rows=2 * eng.gu_half,
                         n_out=d.intermediate_size)
        else:
            eng.forward(ids, pos, B)

    with torch.cuda.stream(s):
        run()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            run()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay (no trace): {e0.elapsed_time(e1) / 40 * 1e3:.1f} us")
    _lib.check(_lib.load().vb_set_trace(buf.data_ptr()), "vb_set_trace")
    g.replay()
    torch.cuda.synchronize()
    _lib.check(_lib.load().vb_set_trace(None), "vb_set_trace")
    h = buf.cpu()
    n = min(int(h[0]), cap)
    recs = h[4:4 + 4 * n].view(n, 4).tolist()
    recs.sort(key=lambda r: r[0])
    t_first = recs[0][0]
    t_end = max(r[1] for r in recs)
    print(f"mode {mode} kv_len {kv_len}: {n} records, step span {(t_end - t_first) / 1e3:.1f} us")
    if mode == "unfused":
        segments(recs, d.num_hidden_layers)
    prev_end = None
    for r in recs[first:first + count]:
        t0, t1, kid, aux = r
        name = NAMES.get(kid, str(kid))
        dur = (t1 - t0) / 1e3
        gap = "" if prev_end is None or name.startswith(" ") else f"  (starts {(t0 - prev_end) / 1e3:+.1f} us vs prev end)"
        print(f"{(t0 - t_first) / 1e3:9.1f} us  {name:22s} aux {aux:2d}  dur {dur:7.1f} us{gap}")
        if not name.startswith(" "):
            prev_end = t1


    fine = h[4 + 4 * cap:].view(10, 256)
    if int(fine.max()) > 0 and len(sys.argv) > 5:
        roles = ["w-issue", "converted", "x-landed", "mma-ready", "committed", "cfull-arrive", "fenced", "x-issued", "mma-b-ready", "-"]
        base = int(fine[fine > 0].min())
        print("fine marks of the last launch, block 0 (us since first mark): slot index g ->", roles)
        if mode == "attn":
            print("attention tiles of block 0 (last launch): tile -> [producer slot free / issue, tile landed (warp 0), warp 0 done]")
            for gi in range(64):
                if int(fine[:3, gi].max()) == 0:
                    break
                print(f"tile {gi:3d}  " + "  ".join(f"{(int(fine[r, gi]) - base) / 1e3:8.2f}" if int(fine[r, gi]) else "       -" for r in range(3)))
            return


if __name__ == "__main__":
    main()
