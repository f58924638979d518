"""End-to-end parity harness (GPU): drive the B200 worker through the in-process scheduler on a seeded
Orpheus-shaped model, then replay the same schedule on the CPU oracle *teacher-forced with the GPU's ids* and
compare, step by step, the sampled ids (bit-exact unless the oracle's top-1/top-2 margin is a bf16 near-tie) and,
chunk by chunk, the PCM bytes (same tokens, same injected NoiseBlock noise).

Used by tests/test_gpu_e2e.py and __graft_entry__.smoke().  Imports oracle/: test infrastructure only.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import orpheus as oorph, sampler as osampler, snac as osnac, worker as oworker


def build_models(dims, snac_cfg, seed, max_bs, page_size, max_pages, greedy=True, lm_head_scale=8.0, stop_boost=None,
                 planted=None):
    """stop_boost: scale of the stop id's lm_head row; with it the stop id is NOT masked (it wins the greedy argmax
    every ~10-20 steps), so the stop / trim / release paths run (orpheus.py:456-466, cuda_graph_worker.py:1176-1277).
    planted: "confident-model" weights (oracle.orpheus.synth_weights): top-1/top-2 margins far above bf16 noise, so the
    ids must match with ZERO mismatches."""
    from vox_serve_b200.engine import LlamaDims
    from vox_serve_b200.model.orpheus import OrpheusModel
    from vox_serve_b200.sampling import SamplingConfig
    from vox_serve_b200.tokenizer.snac import SNAC
    from vox_serve_b200.worker import ModelWorker

    weights = oorph.synth_weights(dims, seed=seed, lm_head_scale=lm_head_scale, planted=planted)
    if stop_boost is not None:
        weights["lm_head.weight"][dims.stop_token_id] *= stop_boost
    snac_sd = osnac.synth_state_dict(snac_cfg, seed=seed + 1)
    ld = LlamaDims(dims.hidden_size, dims.num_hidden_layers, dims.num_attention_heads, dims.num_key_value_heads,
                   dims.head_dim, dims.intermediate_size, dims.vocab_size, dims.rms_norm_eps, dims.rope_theta,
                   dims.rope_factor, dims.low_freq_factor, dims.high_freq_factor, dims.old_context_len)
    snac = SNAC(sampling_rate=snac_cfg.sampling_rate, encoder_dim=snac_cfg.encoder_dim,
                encoder_rates=snac_cfg.encoder_rates, latent_dim=snac_cfg.latent_dim, decoder_dim=snac_cfg.decoder_dim,
                decoder_rates=snac_cfg.decoder_rates, codebook_size=snac_cfg.codebook_size,
                codebook_dim=snac_cfg.codebook_dim, vq_strides=snac_cfg.vq_strides, device="cuda")
    snac.load_state_dict(snac_sd)
    model = OrpheusModel("orpheus-test", device=f"cuda:{torch.cuda.current_device()}", state_dict=weights, dims=ld,
                         snac=snac, stop_token_id=dims.stop_token_id,
                         audio_id_base=dims.audio_id_base, max_tokens=dims.max_tokens, mask_stop_token=stop_boost is None)
    model.default_sampling_config = SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3,
                                                   repetition_window=-1, greedy=greedy, max_tokens=dims.max_tokens)
    worker = ModelWorker("orpheus-test", max_batch_size=max_bs, max_num_pages=max_pages, page_size=page_size,
                         model=model, max_prefill_tokens=256)
    ocfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1,
                                   greedy=greedy, max_tokens=dims.max_tokens)
    ow = oworker.OracleWorker(weights, dims, ocfg, page_size=page_size, max_num_pages=max_pages, snac_sd=snac_sd,
                              snac_cfg=snac_cfg, max_batch_size=max_bs, ignore_stop=stop_boost is None)
    return worker, ow


def run_e2e_parity(prompt_lens=(5, 16, 30, 33), n_tokens=40, seed=3, dims=None, page_size=16, max_pages=128,
                   noise_seed=1234, stop_boost=None, planted=None):
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    dims = dims or oorph.OrpheusDims.tiny()
    dims.max_tokens = max(prompt_lens) + n_tokens
    snac_cfg = osnac.SnacConfig.tiny()
    max_bs = len(prompt_lens)
    worker, ow = build_models(dims, snac_cfg, seed, max_bs, page_size, max_pages, stop_boost=stop_boost, planted=planted)
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist() for n in prompt_lens]

    # ---- GPU run through the worker API ----
    gpu_noise = torch.Generator().manual_seed(noise_seed)
    worker.model.audio_decoder.noise_source = lambda shapes: [torch.randn(s, generator=gpu_noise).cuda() for s in shapes]
    sched = Scheduler(worker)
    sched.trace = []
    reqs = [Request(request_id=f"r{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(prompts)]
    for r in reqs:
        sched.submit(r)
    n_steps = sched.run_until_done(max_steps=4000)
    torch.cuda.synchronize()

    # ---- oracle replay, teacher-forced ----
    ow.noise_gen = torch.Generator().manual_seed(noise_seed)
    oreqs = [oworker.Req(f"r{i}", oworker.format_prompt(p)) for i, p in enumerate(prompts)]
    stats = dict(steps=n_steps, rows=0, id_mismatch=0, low_margin=0, min_margin=1e9, chunks=0, pcm_max_lsb=0,
                 pcm_bytes=0, gpu_launches=worker.gpu_launches)
    active = list(oreqs)
    for step in range(n_steps):
        active = [r for r in active if not r.done_all]
        det = ow.select_detokenize(active)
        lm = ow.select_lm(active, prefill_graph_batch_size=max_bs)
        inp = ow.prepare_lm_inputs(lm)
        ow.run_detokenize(det)
        for r in det:
            if r.done_all:
                ow.free_kv_cache(r)
        if not lm:
            assert step >= len(sched.trace) or not sched.trace[step], (step, sched.trace[step])
            continue
        tr = sched.trace[step]
        assert [x[0] for x in tr] == [r.request_id for r in lm], (step, tr, [r.request_id for r in lm])
        forced = torch.tensor([[x[1]] for x in tr], dtype=torch.int64)
        ow.run_lm(lm, inp, forced_ids=forced)
        pen = ow.last_penalised[:, 0].float()
        top2 = torch.topk(pen, 2, dim=-1).values
        margin = top2[:, 0] - top2[:, 1]
        ulp = top2[:, 0].abs() * 2.0 ** -8
        for i in range(len(lm)):
            stats["rows"] += 1
            stats["min_margin"] = min(stats["min_margin"], float(margin[i]))
            stats["min_margin_ulps"] = min(stats.get("min_margin_ulps", 1e9), float(margin[i] / ulp[i]))
            if int(ow.last_own_ids[i, 0]) != int(forced[i, 0]):
                stats["id_mismatch"] += 1
                if margin[i] <= 4 * ulp[i] and int(forced[i, 0]) in torch.topk(pen[i], 3).indices.tolist():
                    stats["low_margin"] += 1
    # ---- audio ----
    for r, o in zip(reqs, oreqs):
        got, ref = sched.audio[r.request_id], o.output_audio
        assert len(got) == len(ref), (r.request_id, len(got), len(ref))
        for a, b in zip(got, ref):
            assert len(a) == len(b), (r.request_id, len(a), len(b))
            da = np.frombuffer(a, dtype=np.int16).astype(np.int32)
            db = np.frombuffer(b, dtype=np.int16).astype(np.int32)
            if len(da):
                stats["pcm_max_lsb"] = max(stats["pcm_max_lsb"], int(np.abs(da - db).max()))
            stats["chunks"] += 1
            stats["pcm_bytes"] += len(a)
        assert r.finish_reason == o.finish_reason, (r.finish_reason, o.finish_reason)
        assert len(r.lm_output_audio_tokens) == len(o.lm_output_audio_tokens)
    stats["audio_seconds"] = sched.audio_seconds()
    stats["finish_reasons"] = [r.finish_reason for r in reqs]
    stats["n_audio_tokens"] = [len(r.lm_output_audio_tokens) for r in reqs]
    # everything a finished request held is back in the pools
    assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == worker.max_batch_size
    return stats
