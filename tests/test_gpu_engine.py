"""LM forward + sampler on the GPU against the CPU oracle, teacher-forced step by step on the tiny
Orpheus-shaped model of the golden run (same weights, same prompts, same page tables).

Bit-exact greedy ids are required wherever the oracle's top-1/top-2 margin exceeds bf16 rounding noise;
logits must agree to bf16 resolution everywhere."""
import os

import numpy as np
import pytest
import torch

from oracle import orpheus as oorph, sampler as osampler, worker as oworker

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _i32(x):
    return torch.tensor(x, dtype=torch.int32, device="cuda")


def _run(dims, page_size, max_pages, prompt_lens, n_steps, seed, max_bs=4, mode="unfused"):
    from vox_serve_b200 import ops
    from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights

    weights = oorph.synth_weights(dims, seed=seed)
    cfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1,
                                  greedy=True, max_tokens=dims.max_tokens)
    ow = oworker.OracleWorker(weights, dims, cfg, page_size=page_size, max_num_pages=max_pages, max_batch_size=max_bs,
                              ignore_stop=True)
    ld = LlamaDims(dims.hidden_size, dims.num_hidden_layers, dims.num_attention_heads, dims.num_key_value_heads,
                   dims.head_dim, dims.intermediate_size, dims.vocab_size, dims.rms_norm_eps, dims.rope_theta,
                   dims.rope_factor, dims.low_freq_factor, dims.high_freq_factor, dims.old_context_len)
    gw = LlamaWeights.from_state_dict(weights, ld)
    kv = torch.zeros(dims.num_hidden_layers, max_pages, 2, page_size, dims.num_key_value_heads, dims.head_dim,
                     dtype=BF, device="cuda")
    eng = LlamaEngine(gw, kv, page_size, max_rows=256)
    eng.force_unfused = mode.startswith("unfused")
    eng.tiled_acts = mode != "unfused-rows"
    assert eng.fused_ok
    g = torch.Generator().manual_seed(21)
    reqs = [oworker.Req(f"r{i}", torch.randint(10, dims.vocab_size, (n,), generator=g)) for i, n in enumerate(prompt_lens)]
    active = []
    stats = dict(steps=0, rows=0, id_mismatch=0, low_margin=0, max_logit_err=0.0, min_margin=1e9)
    for step in range(n_steps):
        if step < len(reqs):
            active.append(reqs[step])
        lm = ow.select_lm(active)
        inp = ow.prepare_lm_inputs(lm)
        if inp is None:
            break
        # ---- GPU: same inputs, same page tables ----
        ids = inp["input_ids"][:, 0].to(torch.int32).cuda()
        pos = inp["position_ids"].cuda()
        R = ids.numel()
        d_indptr, d_indices = _i32(inp["paged_kv_indptr"]), _i32(inp["paged_kv_indices"])
        d_last = _i32(inp["paged_kv_last_page_len"])
        qo = _i32(inp["qo_indptr"]) if inp["is_prefill"] else None
        ops.plan_rows(eng.plan, qo, d_indptr, d_indices, d_last, len(lm), R, page_size, eng.chunk)
        last_rows = _i32([x - 1 for x in inp["qo_indptr"][1:]]) if inp["is_prefill"] else None
        logits = eng.forward(ids, pos, R, last_rows=last_rows)
        rep = inp["repetition_cache"].clone()
        gpu_ids = ops.sample(logits, "greedy", rep_cache=rep.cuda(), penalty=cfg.repetition_penalty,
                             mask_token=dims.stop_token_id).cpu()
        # ---- oracle ----
        ref_ids = ow.run_lm(lm, inp)[:, 0]
        ref_logits = ow.last_logits[:, 0].float()
        err = (logits.float().cpu() - ref_logits).abs().max().item()
        scale = ref_logits.abs().max().item()
        stats["max_logit_err"] = max(stats["max_logit_err"], err / scale)
        pen = osampler.apply_repetition_penalty(ow.last_logits, rep, cfg.repetition_penalty)[:, 0].float()
        pen[:, dims.stop_token_id] = float("-inf")
        top2 = torch.topk(pen, 2, dim=-1).values
        margin = top2[:, 0] - top2[:, 1]
        ulp = top2[:, 0].abs() * 2.0 ** -8
        for r in range(len(lm)):
            stats["rows"] += 1
            stats["min_margin"] = min(stats["min_margin"], float(margin[r]))
            if int(gpu_ids[r]) != int(ref_ids[r]):
                stats["id_mismatch"] += 1
                # only a near-tie may flip, and then only to the runner-up
                assert margin[r] <= 4 * ulp[r], (step, r, float(margin[r]), float(ulp[r]))
                stats["low_margin"] += 1
                assert int(gpu_ids[r]) in torch.topk(pen[r], 3).indices.tolist()
        stats["steps"] += 1
    return stats


# the default 8-launch layer (tiled / row-major activations) and the fused projections (5 launches per layer)
@pytest.mark.parametrize("mode", ["unfused", "unfused-rows", "fused"])
def test_tiny_orpheus_teacher_forced_greedy(mode):
    dims = oorph.OrpheusDims.tiny()
    dims.max_tokens = 400
    st = _run(dims, page_size=16, max_pages=128, prompt_lens=[5, 16, 30, 33], n_steps=60, seed=3,
              mode=mode)
    print(st)
    assert st["max_logit_err"] < 2e-2
    assert st["id_mismatch"] <= max(2, st["rows"] // 50)


@pytest.mark.parametrize("mode", ["unfused", "fused"])
def test_medium_orpheus_page128_teacher_forced(mode):
    # head_dim 128, GQA 3, page 128: the Orpheus attention geometry with a prompt crossing a page
    dims = oorph.OrpheusDims.tiny(hidden_size=1536, num_hidden_layers=3, num_attention_heads=12,
                                  num_key_value_heads=4, intermediate_size=2048, vocab_size=10 + 7 * 4096)
    dims.max_tokens = 400
    st = _run(dims, page_size=128, max_pages=16, prompt_lens=[133, 120, 7], n_steps=24, seed=4, mode=mode)
    print(st)
    assert st["max_logit_err"] < 2e-2
    assert st["id_mismatch"] <= max(2, st["rows"] // 50)
