"""The BENCHMARKED configuration against the CPU oracle: Orpheus-3B at its true dimensions (28 layers, hidden 3072,
24/8 heads x 128, intermediate 8192, vocab 156 940 -> 1227 lm_head tiles, production split-K and gate/up tiling),
batch 32, page 128, ragged prompts whose KV crosses page boundaries, DEFAULT decode mode, CUDA graphs -- i.e. the
tokens bench.py times (SURVEY.md §7 "hard parts", vox_serve/model/orpheus.py:125-221, worker/base.py:299).

* free-running greedy (north_star: "audio-token IDs bit-exact under greedy decode"): the B200 worker and the oracle
  worker each run on their OWN tokens for 32 prefill steps + 64 decode steps; ZERO id mismatches are required and the
  smallest top-1/top-2 margin of the oracle's penalised logits is reported (the synthetic weights are the "confident
  model" of oracle.orpheus.synth_weights(planted=...): margins are > 100 bf16 ulps of the top logit, not near-ties);
  the logits of every step are also held to 2e-2 of the row scale.
* teacher-forced with i.i.d. weights (no planted direction: every logit is pure network output), on the engine with
  the production launch plan.  A 28-layer network of i.i.d. N(0, 0.02) weights amplifies rounding noise: two CPU
  evaluations of the SAME oracle arithmetic that differ only in the fp32 summation order of the dot products (K summed
  in one piece / in two halves) disagree by 3-4 % of the row scale and flip ~8 % of the greedy ids (measured, see
  ``_oracle_noise_floor``).  The GPU is therefore held to twice that measured floor, and every id it picks differently
  from the oracle must be a provable near-tie: the oracle's margin between the two candidates is below the logit
  disagreement measured at those two tokens.  (The 2e-2 / bit-exact claims are carried by the test above.)
"""
import time

import pytest
import torch

from oracle import orpheus as oorph, sampler as osampler, snac as osnac, worker as oworker

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _llama_dims(dims):
    from vox_serve_b200.engine import LlamaDims

    return LlamaDims(dims.hidden_size, dims.num_hidden_layers, dims.num_attention_heads, dims.num_key_value_heads,
                     dims.head_dim, dims.intermediate_size, dims.vocab_size, dims.rms_norm_eps, dims.rope_theta,
                     dims.rope_factor, dims.low_freq_factor, dims.high_freq_factor, dims.old_context_len)


def test_orpheus_3b_true_dims_free_running_greedy_bit_exact():
    from vox_serve_b200.model.orpheus import OrpheusModel
    from vox_serve_b200.requests import Request
    from vox_serve_b200.sampling import SamplingConfig
    from vox_serve_b200.scheduler import Scheduler
    from vox_serve_b200.tokenizer.snac import SNAC
    from vox_serve_b200.worker import ModelWorker

    n_req, n_decode, page, pages = 32, 64, 128, 32 * 3
    dims = oorph.OrpheusDims()
    dims.max_tokens = 1200
    t0 = time.time()
    weights = oorph.synth_weights(dims, seed=11, planted=2.0)
    t_w = time.time() - t0
    scfg = osnac.SnacConfig.tiny()
    snac = SNAC(sampling_rate=scfg.sampling_rate, encoder_dim=scfg.encoder_dim, encoder_rates=scfg.encoder_rates,
                latent_dim=scfg.latent_dim, decoder_dim=scfg.decoder_dim, decoder_rates=scfg.decoder_rates,
                codebook_size=scfg.codebook_size, codebook_dim=scfg.codebook_dim, vq_strides=scfg.vq_strides, device="cuda")
    snac.load_state_dict(osnac.synth_state_dict(scfg, seed=12))
    model = OrpheusModel("orpheus-test", state_dict=weights, dims=_llama_dims(dims), snac=snac, max_tokens=dims.max_tokens,
                         mask_stop_token=True)
    kw = dict(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1, greedy=True, max_tokens=dims.max_tokens)
    model.default_sampling_config = SamplingConfig(**kw)
    worker = ModelWorker("orpheus-test", max_batch_size=n_req, max_num_pages=pages, page_size=page, model=model,
                         max_prefill_tokens=256)
    eng = model.engine_for(worker.kv_cache, page)
    assert eng.force_unfused, "the default decode mode is what bench.py times"
    assert (eng.split_qkv, eng.split_o, eng.split_down) == (4, 4, 8), "production split-K of Orpheus-3B at 32 rows"
    ow = oworker.OracleWorker(weights, dims, osampler.SamplingConfig(**kw), page_size=page, max_num_pages=pages,
                              max_batch_size=n_req, ignore_stop=True)
    g = torch.Generator().manual_seed(5)
    lens = [90 + int(torch.randint(0, 110, (1,), generator=g)) for _ in range(n_req)]     # 90..199: pages 1 -> 2 -> 3
    prompts = [torch.randint(0, 128000, (n - 5,), generator=g).tolist() for n in lens]
    sched = Scheduler(worker)
    sched.trace = []
    reqs = [Request(request_id=f"r{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(prompts)]
    oreqs = [oworker.Req(f"r{i}", oworker.format_prompt(p)) for i, p in enumerate(prompts)]
    for r in reqs:
        sched.submit(r)
    st = dict(rows=0, id_mismatch=0, min_rel_margin=1e9, max_logit_err=0.0, graph_steps=0, t_weights=round(t_w, 1))
    active = list(oreqs)
    t0 = time.time()
    for step in range(n_req + n_decode):
        # ---- B200 worker, on its own tokens ----
        n_lm, _ = sched._step()
        torch.cuda.synchronize()
        gpu_logits = eng.logits[:n_lm].float().cpu()
        tr = sched.trace[step]
        st["graph_steps"] += int(step >= n_req)
        # ---- oracle worker, on ITS own tokens (no teacher forcing) ----
        lm = ow.select_lm(active, prefill_graph_batch_size=n_req)
        assert [x[0] for x in tr] == [r.request_id for r in lm]
        inp = ow.prepare_lm_inputs(lm)
        rep = inp["repetition_cache"].clone()
        ids = ow.run_lm(lm, inp)
        ref_logits = ow.last_logits[:, 0].float()
        pen = osampler.apply_repetition_penalty(ow.last_logits, rep, 1.3)[:, 0].float()
        pen[:, dims.stop_token_id] = float("-inf")
        top2 = torch.topk(pen, 2, dim=-1).values
        st["min_rel_margin"] = min(st["min_rel_margin"], float(((top2[:, 0] - top2[:, 1]) / top2[:, 0].abs()).min()))
        st["max_logit_err"] = max(st["max_logit_err"],
                                  float((gpu_logits - ref_logits).abs().max() / ref_logits.abs().max()))
        for i in range(len(lm)):
            st["rows"] += 1
            st["id_mismatch"] += int(tr[i][1] != int(ids[i, 0]))
        assert st["id_mismatch"] == 0, (step, st)       # free-running: a fork would invalidate everything after it
    st["seconds"] = round(time.time() - t0, 1)
    st["min_margin_ulps"] = round(st["min_rel_margin"] * 256, 1)
    st["kv_len_range"] = (min(r.kv_token_len for r in reqs), max(r.kv_token_len for r in reqs))
    print("true-dims free-running:", st)
    assert st["graph_steps"] == n_decode and B32_graph_captured(worker, n_req)
    assert st["id_mismatch"] == 0 and st["rows"] >= n_req * n_decode
    assert st["min_rel_margin"] * 256 > 16, st          # >> 4 bf16 ulps: no near-tie was involved in the match
    assert st["max_logit_err"] < 2e-2, st
    assert st["kv_len_range"][1] > 2 * page            # some rows are on their third page


def B32_graph_captured(worker, n):
    return n in worker.decode_graphs


def _oracle_noise_floor(ow, dims, inp, page):
    """Relative logit disagreement (max |a - b| / max |a| per row, worst row) and id-flip fraction between two CPU
    evaluations of oracle.orpheus.lm_forward on one decode step: torch's bf16 matmul against the same products summed
    in fp32 over the two halves of K.  Both are the reference's arithmetic (nn.Linear in bf16 with fp32 accumulation)."""
    import torch.nn.functional as F

    from oracle import lm_ops

    def split_k_linear(x, wt, b=None):
        h = x.shape[-1] // 2
        return ((x[..., :h].float() @ wt[:, :h].float().t()) + (x[..., h:].float() @ wt[:, h:].float().t())).to(x.dtype)

    outs = []
    for lin in (F.linear, split_k_linear):
        kv = ow.kv_cache.clone()
        wr = lm_ops.PagedWrapperCPU("decode", page)
        wr.plan(inp["paged_kv_indptr"], inp["paged_kv_indices"], inp["paged_kv_last_page_len"])
        keep, F.linear = F.linear, lin
        try:
            outs.append(oorph.lm_forward(ow.w, dims, inp["input_ids"][:, 0], inp["position_ids"], wr, kv).float())
        finally:
            F.linear = keep
    a, b = outs
    rel = float(((a - b).abs().amax(-1) / a.abs().amax(-1)).max())
    return rel, float((a.argmax(-1) != b.argmax(-1)).float().mean())


def test_orpheus_3b_true_dims_teacher_forced_iid_weights():
    from vox_serve_b200 import ops
    from vox_serve_b200.engine import LlamaEngine, LlamaWeights

    n_req, n_decode, page, pages = 32, 4, 128, 32 * 2
    dims = oorph.OrpheusDims()
    dims.max_tokens = 1200
    weights = oorph.synth_weights(dims, seed=7)
    cfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1, greedy=True,
                                  max_tokens=dims.max_tokens)
    ow = oworker.OracleWorker(weights, dims, cfg, page_size=page, max_num_pages=pages, max_batch_size=n_req, ignore_stop=True)
    gw = LlamaWeights.from_state_dict(weights, _llama_dims(dims))
    kv = torch.zeros(dims.num_hidden_layers, pages, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF, device="cuda")
    eng = LlamaEngine(gw, kv, page, max_rows=256 + n_req)
    assert eng.force_unfused and (eng.split_qkv, eng.split_o, eng.split_down) == (4, 4, 8)
    g = torch.Generator().manual_seed(9)
    lens = [100 + int(torch.randint(0, 50, (1,), generator=g)) for _ in range(n_req)]
    lens[3], lens[17] = 128, 127           # last_page_len == page_size right after prefill / after the first decode step
    reqs = [oworker.Req(f"r{i}", torch.randint(0, 128000, (n,), generator=g)) for i, n in enumerate(lens)]
    active = list(reqs)
    st = dict(rows=0, id_mismatch=0, low_margin=0, max_logit_err=0.0)

    def i32(x):
        return torch.tensor(x, dtype=torch.int32, device="cuda")

    for step in range(n_req + n_decode):
        lm = ow.select_lm(active, prefill_graph_batch_size=n_req)
        inp = ow.prepare_lm_inputs(lm)
        if step == n_req + n_decode - 1:
            st["oracle_floor"], st["oracle_flip_frac"] = _oracle_noise_floor(ow, dims, inp, page)
        ids = inp["input_ids"][:, 0].to(torch.int32).cuda()
        R = ids.numel()
        qo = i32(inp["qo_indptr"]) if inp["is_prefill"] else None
        ops.plan_rows(eng.plan, qo, i32(inp["paged_kv_indptr"]), i32(inp["paged_kv_indices"]),
                      i32(inp["paged_kv_last_page_len"]), len(lm), R, page, eng.chunk)
        last_rows = i32([x - 1 for x in inp["qo_indptr"][1:]]) if inp["is_prefill"] else None
        logits = eng.forward(ids, inp["position_ids"].cuda(), R, last_rows=last_rows)
        rep = inp["repetition_cache"].clone()
        gpu_ids = ops.sample(logits, "greedy", rep_cache=rep.cuda(), penalty=cfg.repetition_penalty,
                             mask_token=dims.stop_token_id).cpu()
        ow.run_lm(lm, inp, forced_ids=gpu_ids.view(-1, 1))
        ref_logits = ow.last_logits[:, 0].float()
        st["max_logit_err"] = max(st["max_logit_err"],
                                  float((logits.float().cpu() - ref_logits).abs().max() / ref_logits.abs().max()))
        pen = ow.last_penalised[:, 0].float()
        got = logits.float().cpu()
        for r in range(len(lm)):
            st["rows"] += 1
            g_id, o_id = int(gpu_ids[r]), int(ow.last_own_ids[r, 0])
            if g_id != o_id:
                # The two implementations may only disagree where the oracle's own margin between the two candidates
                # is smaller than the (tolerated, measured) disagreement of their logits at those two tokens: 28 layers
                # of i.i.d. weights put ~1e-2 of the row scale of rounding noise on every logit.  The repetition penalty
                # scales a logit by at most 1.3 either way.
                st["id_mismatch"] += 1
                e = float((got[r, g_id] - ref_logits[r, g_id]).abs() + (got[r, o_id] - ref_logits[r, o_id]).abs())
                margin = float(pen[r, o_id] - pen[r, g_id])
                assert 0 <= margin <= 1.3 * e + 1e-6, (step, r, margin, e)
                # (the two deviations are each bounded by the measured noise floor below, once it is known)
                st["max_pair_err"] = max(st.get("max_pair_err", 0.0), e / float(ref_logits[r].abs().max()))
                st["low_margin"] += 1
                st["max_flip_margin"] = max(st.get("max_flip_margin", 0.0), margin)
    print("true-dims teacher-forced:", st)
    assert st["oracle_floor"] > 5e-3, st                 # (if this ever drops, tighten the bound below to 2e-2)
    assert st["max_logit_err"] < 2 * st["oracle_floor"], st
    assert st.get("max_pair_err", 0.0) <= 2 * 2 * st["oracle_floor"], st
    assert st["id_mismatch"] <= 3 * max(st["oracle_flip_frac"], 0.05) * st["rows"], st
