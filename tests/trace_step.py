"""Dev tool (GPU): in-situ timeline of one CUDA-graph replay of the decode step (block 0 of every instrumented kernel
logs %globaltimer through vb_set_trace).  python tests/trace_step.py [kv_len] [mode: chain|fused|unfused] [first] [count]"""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import _lib, ops  # noqa: E402
from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights  # noqa: E402
from vox_serve_b200.model.orpheus import synthetic_state_dict  # noqa: E402

NAMES = {5: "reduce+norm", 6: "rope+append", 24: " rope:dep-released", 22: " attn:first-tile", 23: " attn:last-tile", 1: "gemm", 2: "attn", 3: "chain", 4: "sample", 10: "rope_tab", 20: " gemm:dep-released", 21: " gemm:acc-done"}
for p in range(4):
    NAMES[30 + p] = f" chain:x-ready p{p}"
    NAMES[40 + p] = f" chain:acc-done p{p}"
    NAMES[50 + p] = f" chain:epi-done p{p}"


def main():
    kv_len = int(sys.argv[1]) if len(sys.argv) > 1 else 728
    mode = sys.argv[2] if len(sys.argv) > 2 else "chain"
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
    eng.use_chain = mode == "chain"
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
        print("fine marks of the LAST chain launch, block 0 (us since first mark): slot index g ->", roles)
        if mode == "attn":
            print("attention tiles of block 0 (last launch): tile -> [producer slot free / issue, tile landed (warp 0), warp 0 done]")
            for gi in range(64):
                if int(fine[:3, gi].max()) == 0:
                    break
                print(f"tile {gi:3d}  " + "  ".join(f"{(int(fine[r, gi]) - base) / 1e3:8.2f}" if int(fine[r, gi]) else "       -" for r in range(3)))
            return
        print("epilogue marks per phase [acc-done, partial-stored, released, peers-in, tail-done, grid-arrived, next-dep-seen]:")
        for ph in range(4):
            row = [int(fine[8, ph * 8 + i]) for i in range(7)]
            print(f"  p{ph}: " + "  ".join(f"{(v - base) / 1e3:8.2f}" if v else "       -" for v in row))
        for q in range(4):
            print(f"epilogue marks, TMEM quarter {q} [start, ld0, bar, store-done, bar, data-seen | +4: chunk 1]:",
                  [round((int(v) - base) / 1e3, 2) if int(v) else None for v in fine[6 + q, :14]])
        for gi in range(256):
            if int(fine[:5, gi].max()) == 0:
                break
            print(f"g {gi:3d}  " + "  ".join(f"{(int(fine[r, gi]) - base) / 1e3:8.2f}" if int(fine[r, gi]) else "       -" for r in range(8)))


if __name__ == "__main__":
    main()
