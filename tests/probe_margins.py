"""Dev tool (CPU): free-running greedy decode of the oracle at TRUE Orpheus-3B dims with the planted synthetic
weights; prints the top-1/top-2 margin statistics that tests/test_gpu_true_dims.py relies on.
usage: python tests/probe_margins.py [n_requests] [n_decode_steps] [planted_std]"""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from oracle import orpheus as oorph, sampler as osampler, worker as oworker  # noqa: E402


def main():
    n_req = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    planted = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    torch.set_num_threads(8)
    dims = oorph.OrpheusDims()
    t0 = time.time()
    w = oorph.synth_weights(dims, seed=11, planted=planted if planted > 0 else None)
    print(f"weights {time.time() - t0:.1f} s", flush=True)
    cfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1, greedy=True,
                                  max_tokens=1200)
    ow = oworker.OracleWorker(w, dims, cfg, page_size=128, max_num_pages=4 * n_req, max_batch_size=n_req, ignore_stop=True)
    g = torch.Generator().manual_seed(5)
    lens = [100 + int(torch.randint(0, 60, (1,), generator=g)) for _ in range(n_req)]
    reqs = [oworker.Req(f"r{i}", oworker.format_prompt(torch.randint(0, 128000, (n - 5,), generator=g).tolist()))
            for i, n in enumerate(lens)]
    active = list(reqs)
    rel = []
    for step in range(n_req + n_steps):
        lm = ow.select_lm(active, prefill_graph_batch_size=n_req)
        inp = ow.prepare_lm_inputs(lm)
        t1 = time.time()
        rep = inp["repetition_cache"].clone()
        ow.run_lm(lm, inp)
        pen = osampler.apply_repetition_penalty(ow.last_logits, rep, cfg.repetition_penalty)[:, 0].float()
        pen[:, dims.stop_token_id] = float("-inf")
        top2 = torch.topk(pen, 2, dim=-1).values
        m = ((top2[:, 0] - top2[:, 1]) / top2[:, 0].abs()).tolist()
        rel.extend(m)
        print(f"step {step} rows {len(lm)} {time.time() - t1:.2f} s  top1 {top2[:, 0].min():.2f}..{top2[:, 0].max():.2f} "
              f"rel margin min {min(m):.4f} (in ulps of the top logit: {min(m) * 256:.1f})", flush=True)
    print(f"ALL: rel margin min {min(rel):.4f} = {min(rel) * 256:.1f} bf16 ulps of the top logit over {len(rel)} row-steps")


if __name__ == "__main__":
    main()
