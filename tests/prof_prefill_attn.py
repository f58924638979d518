"""Timing probe (GPU): prefill-shaped attention steps on the tiled tensor-core kernel (vb_paged_prefill_attn) and on
the one-stream-per-row kernel (vb_paged_attn), same plan, same K/V.  One JSON line per scenario:

    python tests/prof_prefill_attn.py > gpurun_out/prefill_attn.jsonl

Not a bench value: a per-kernel probe for DESIGN.md / profiles/.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from vox_serve_b200 import ops  # noqa: E402

SCENARIOS = [
    # name, new tokens per request, kv length per request, page, q heads, kv heads, head_dim
    ("orpheus 1 x 133-token prompt", [133], [133], 128, 24, 8, 128),
    ("orpheus 8 x 133-token prompts", [133] * 8, [133] * 8, 128, 24, 8, 128),
    ("orpheus 133-token prompt joining 31 decodes (kv 193..763)", [1] * 31 + [133],
     [193 + 19 * i for i in range(31)] + [133], 128, 24, 8, 128),
    ("glm-4-voice 1 x 435-token prompt", [435], [435], 128, 32, 2, 128),
    ("csm backbone 4 x 600-row prompts", [600] * 4, [600] * 4, 32, 32, 8, 64),
    ("cosyvoice2 1 x 300-row prompt", [300], [300], 128, 14, 2, 64),
    ("qwen3-tts talker 16 x 50-row prompts", [50] * 16, [50] * 16, 128, 16, 8, 128),
]


def main():
    dev = "cuda"
    for name, new, kv, page, hq, hkv, D in SCENARIOS:
        n_req = len(new)
        n_pages = sum((L + page - 1) // page for L in kv) + 2
        perm = torch.randperm(n_pages).tolist()
        indptr, indices, last = [0], [], []
        for L in kv:
            n = (L + page - 1) // page
            indices += [perm.pop() for _ in range(n)]
            indptr.append(len(indices))
            last.append(L - (n - 1) * page)
        qo = [0] + [int(x) for x in np.cumsum(new)]
        R = qo[-1]
        cache = torch.randn(1, n_pages, 2, page, hkv, D, device=dev).to(torch.bfloat16)
        q = torch.randn(R, hq, D, device=dev).to(torch.bfloat16)
        chunk = ops.attn_chunk_tokens(page, hkv)
        plan = ops.RowPlan(max(R, 8), dev)
        i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)  # noqa: E731
        ops.plan_rows(plan, i32(qo), i32(indptr), i32(indices), i32(last), n_req, R, page, chunk)
        ws = ops.paged_attn_workspace(R, None, hq, hkv, D, dev)
        out = torch.empty_like(q)
        res = {"scenario": name, "rows": R, "requests": n_req,
               "kv_bytes_once": int(sum(kv) * hkv * D * 2 * 2),
               "auto_choice": "tiles" if ops.use_prefill_tiles(plan, R, D, page) else "rows"}
        outs = {}
        for label, tiles in (("tiles", True), ("tc", "tc"), ("rows", False)):
            def run():
                ops.paged_attn(q, cache, 0, plan, R, hkv, page, chunk, ws, out=out, prefill_tiles=tiles)
            for _ in range(5):
                run()
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                with torch.cuda.graph(gph, stream=s):
                    for _ in range(28):
                        run()
            torch.cuda.current_stream().wait_stream(s)
            gph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                gph.replay()
            e1.record()
            torch.cuda.synchronize()
            res[f"{label}_us_per_launch"] = round(e0.elapsed_time(e1) * 1e3 / 280, 2)
            outs[label] = out.float().clone()
        res["max_abs_diff"] = float((outs["tiles"] - outs["rows"]).abs().max())
        res["max_abs_diff_tc"] = float((outs["tc"] - outs["rows"]).abs().max())
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
