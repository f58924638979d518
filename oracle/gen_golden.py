"""Generate tests/golden/*.npz by executing the REFERENCE's own code on CPU (authoring container only).

    python -m oracle.gen_golden            # writes tests/golden/

What runs from /root/reference, unmodified (see oracle/ref_import.py for the import recipe):
  * ``OrpheusForCausalLM`` / ``OrpheusModel.forward|sampling|postprocess`` (model/orpheus.py)
  * ``Sampler`` (sampling.py)                      * ``SNAC`` decoder (tokenizer/snac.py)
  * ``ModelWorker.prepare_lm_inputs|run_detokenize|free_kv_cache`` (worker/base.py)
  * ``Scheduler._select_lm_requests|_select_detokenize_requests`` (scheduler/base.py)
What is substituted: the three CUDA-only FlashInfer ops (oracle/lm_ops.py restatements) and the
``torch.randn`` NoiseBlock draws (served from a seeded CPU generator so they can be replayed).

Weights are NOT stored: they are re-derived from seeds by ``oracle.orpheus.synth_weights`` /
``oracle.snac.synth_state_dict`` (torch CPU generator; same torch build on the GPU box).
"""
from __future__ import annotations

import asyncio
import os
import queue
import sys
import types

import numpy as np
import torch

from . import lm_ops, orpheus as oorph, sampler as osampler, snac as osnac
from .ref_import import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ------------------------------------------------------------------------------------------
def build_ref_orpheus(ref, dims: oorph.OrpheusDims, weights, snac_cfg, snac_sd, cfg):
    from transformers import LlamaConfig

    hf = LlamaConfig(
        vocab_size=dims.vocab_size, hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
        num_hidden_layers=dims.num_hidden_layers, num_attention_heads=dims.num_attention_heads,
        num_key_value_heads=dims.num_key_value_heads, head_dim=dims.head_dim, rms_norm_eps=dims.rms_norm_eps,
        hidden_act="silu", attention_bias=False, mlp_bias=False, pad_token_id=None, tie_word_embeddings=False,
    )
    hf.rope_theta = dims.rope_theta
    hf.rope_scaling = {"factor": dims.rope_factor, "low_freq_factor": dims.low_freq_factor,
                       "high_freq_factor": dims.high_freq_factor,
                       "original_max_position_embeddings": dims.old_context_len, "rope_type": "llama3"}
    lm = ref.orpheus.OrpheusForCausalLM(hf)
    missing = lm.load_state_dict(weights, strict=True)
    lm = lm.to(torch.bfloat16).eval()

    snac = ref.snac.SNAC(
        sampling_rate=snac_cfg.sampling_rate, encoder_dim=snac_cfg.encoder_dim,
        encoder_rates=list(snac_cfg.encoder_rates), decoder_dim=snac_cfg.decoder_dim,
        decoder_rates=list(snac_cfg.decoder_rates), attn_window_size=None,
        codebook_size=snac_cfg.codebook_size, codebook_dim=snac_cfg.codebook_dim,
        vq_strides=list(snac_cfg.vq_strides), noise=True, depthwise=True).eval()
    r = snac.load_state_dict(snac_sd, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all(k.startswith("encoder") or ".in_proj" in k for k in r.missing_keys), r.missing_keys

    m = object.__new__(ref.orpheus.OrpheusModel)     # skip from_pretrained (orpheus.py:238-250)
    m.model_name, m.device, m.dtype = "synthetic-orpheus", "cpu", torch.bfloat16
    m.enable_torch_compile, m.audio_decoder_device = False, "cpu"
    m.model, m.audio_decoder = lm, snac
    m._num_attention_heads, m._num_key_value_heads = dims.num_attention_heads, dims.num_key_value_heads
    m._num_hidden_layers, m._hidden_size = dims.num_hidden_layers, dims.hidden_size
    m.stop_token_id = dims.stop_token_id
    m.default_sampling_config = cfg
    m.idx_14 = torch.tensor([1, 4], dtype=torch.long)
    m.idx_2356 = torch.tensor([2, 3, 5, 6], dtype=torch.long)
    # audio id base differs from 128266 for the tiny vocab: same formula, other constant
    base = dims.audio_id_base
    m._turn_token_into_id = lambda ids: (ids - base) % 4096
    m.preprocess = types.MethodType(_preprocess_ids, m)
    type(m).max_tokens  # property exists
    return m


def _preprocess_ids(self, prompt=None, audio_path=None, **kw):
    """Token ids arrive pre-tokenised (no HF tokenizer on disk): same output structure as
    OrpheusModel.preprocess (orpheus.py:366-396)."""
    from vox_serve.model.base import PreprocessOutput

    ids = torch.as_tensor(prompt, dtype=torch.int64).view(-1, 1)
    cfg = self.default_sampling_config
    rep = None
    if cfg.repetition_penalty is not None and cfg.repetition_window is not None and cfg.repetition_penalty != 1.0:
        rep = torch.zeros(cfg.repetition_window if cfg.repetition_window > 0 else 1, 1, self.vocab_size,
                          dtype=torch.bool)
    return PreprocessOutput(input_tokens=ids, repetition_cache=rep)


def build_ref_worker(ref, model, page_size, max_num_pages, max_batch_size):
    w = object.__new__(ref.graph_worker.CudaGraphWorker)   # skip CUDA init (worker/base.py:15-125)
    w.model, w.device, w.detokenizer_device = model, "cpu", "cpu"
    w.max_batch_size, w.page_size, w.max_num_pages = max_batch_size, page_size, max_num_pages
    w.empty_pages = queue.Queue()
    for i in range(max_num_pages):
        w.empty_pages.put(i)
    w.needs_watermarking, w.nvtx_enabled = False, False
    import logging
    w.logger = logging.getLogger("golden")
    w.prefill_graph_batch_size, w.cuda_graph_seq_len_buckets = 8, [1024]
    w.kv_cache = torch.zeros(model.num_hidden_layers, max_num_pages, 2, page_size, model.num_key_value_heads,
                             model.head_dim, dtype=torch.bfloat16)
    return w


def ref_lm_step(ref, worker, reqs, lm_inputs, stop_mask_id=None):
    """plan -> model.forward -> last-token gather (cuda_graph_worker.py:900-902) -> model.sampling
    -> update_req_states, all reference code except the paged wrapper."""
    if not reqs:
        return None
    ps = worker.page_size
    if lm_inputs["is_prefill"]:
        wr = lm_ops.PagedWrapperCPU("prefill", ps)
        wr.plan(lm_inputs["qo_indptr"], lm_inputs["paged_kv_indptr"], lm_inputs["paged_kv_indices"],
                lm_inputs["paged_kv_last_page_len"])
    else:
        wr = lm_ops.PagedWrapperCPU("decode", ps)
        wr.plan(lm_inputs["paged_kv_indptr"], lm_inputs["paged_kv_indices"], lm_inputs["paged_kv_last_page_len"])
    logits = worker.model.forward(input_ids=lm_inputs["input_ids"], position_ids=lm_inputs["position_ids"],
                                  attn_wrapper=wr, kv_cache=worker.kv_cache)
    if lm_inputs["is_prefill"]:
        logits = logits[torch.tensor(lm_inputs["qo_indptr"][1:]) - 1]
    raw = logits.clone()
    if stop_mask_id is not None:
        logits = logits.clone()
        logits[..., stop_mask_id] = float("-inf")
    ids, task = worker.model.sampling(logits=logits, requests=reqs, repetition_cache=lm_inputs["repetition_cache"])
    asyncio.run(task)
    return raw, ids


class _NoisePatch:
    """Serve NoiseBlock's torch.randn((B,1,T), device=, dtype=) from a seeded CPU generator."""

    def __init__(self, seed):
        self.gen = torch.Generator().manual_seed(seed)
        self.log = []

    def __enter__(self):
        self._orig = torch.randn

        def fake(*size, **kw):
            shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list)) else size
            t = self._orig(tuple(shape), generator=self.gen)
            self.log.append(t)
            return t.to(kw.get("dtype", torch.float32))

        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self._orig


# ------------------------------------------------------------------------------------------
def golden_sampler(ref):
    g = torch.Generator().manual_seed(11)
    S = ref.sampling.Sampler
    out = {}
    B, V = 5, 97
    logits = (torch.randn(B, 1, V, generator=g) * 3).to(torch.bfloat16)
    cache = torch.rand(B, 1, 1, V, generator=g) < 0.3
    out["pen_logits"], out["pen_cache"] = logits.float().numpy(), cache.numpy()
    out["pen_out"] = S.apply_repetition_penalty(logits, cache, 1.3).float().numpy()
    ids = torch.argmax(torch.from_numpy(out["pen_out"]), -1)
    out["greedy_ids"] = S.run_sampling(torch.from_numpy(out["pen_out"]).to(torch.bfloat16).view(-1, V),
                                       ref.sampling.SamplingConfig(greedy=True)).numpy()
    c2 = cache.clone()
    S.update_repetition_penalty_cache(c2, ids, -1)
    out["upd_global_ids"], out["upd_global_out"] = ids.numpy(), c2.numpy()
    # windowed, multi-codebook
    W, C = 3, 2
    cw = torch.rand(B, W, C, V, generator=g) < 0.2
    idw = torch.randint(0, V, (B, C), generator=g)
    out["upd_win_in"], out["upd_win_ids"] = cw.numpy(), idw.numpy()
    c3 = cw.clone()
    S.update_repetition_penalty_cache(c3, idw, W)
    out["upd_win_out"] = c3.numpy()
    lw = (torch.randn(B, C, V, generator=g) * 3).to(torch.bfloat16)
    out["pen_win_logits"] = lw.float().numpy()
    out["pen_win_out"] = S.apply_repetition_penalty(lw, cw, 1.7).float().numpy()
    # codebook-0-only logits against a multi-codebook cache (sampling.py:140-141, 167-175)
    l0 = lw[:, :1]
    out["pen_cb0_out"] = S.apply_repetition_penalty(l0, cw, 1.7).float().numpy()
    c4 = cw.clone()
    S.update_repetition_penalty_cache(c4, idw[:, :1], W)
    out["upd_cb0_win_out"] = c4.numpy()
    c5 = cw.clone()
    S.update_repetition_penalty_cache(c5, idw[:, :1], -1)
    out["upd_cb0_global_out"] = c5.numpy()
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **out)
    print("sampler.npz", {k: v.shape for k, v in out.items()})


def golden_snac(ref):
    for tag, cfg, B in (("tiny", osnac.SnacConfig.tiny(), 3), ("24khz", osnac.SnacConfig(), 2)):
        sd = osnac.synth_state_dict(cfg, seed=5)
        m = ref.snac.SNAC(
            sampling_rate=cfg.sampling_rate, encoder_dim=cfg.encoder_dim, encoder_rates=list(cfg.encoder_rates),
            decoder_dim=cfg.decoder_dim, decoder_rates=list(cfg.decoder_rates), attn_window_size=None,
            codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim, vq_strides=list(cfg.vq_strides),
            noise=True, depthwise=True).eval()
        r = m.load_state_dict(sd, strict=False)
        assert not r.unexpected_keys
        g = torch.Generator().manual_seed(6)
        nf = 4
        codes = [torch.randint(0, cfg.codebook_size, (B, nf * cfg.vq_strides[0] // s), generator=g)
                 for s in cfg.vq_strides]
        with _NoisePatch(77) as npatch, torch.inference_mode():
            wav = m.decode(codes)
        d = {f"codes{i}": c.numpy() for i, c in enumerate(codes)}
        d.update({f"noise{i}": n.numpy() for i, n in enumerate(npatch.log)})
        d["wav"] = wav.numpy()
        np.savez_compressed(os.path.join(OUT, f"snac_{tag}.npz"), **d)
        print(f"snac_{tag}.npz", wav.shape, float(wav.abs().mean()), float(wav.abs().max()))


def golden_orpheus_e2e(ref):
    """Tiny Orpheus through the reference's scheduler-selection + worker + adapter code.

    5 requests with ragged prompts (one exactly a page, one crossing a page boundary mid-decode),
    page_size 16, greedy + repetition penalty 1.3 / window -1 (the parity configuration,
    model/__init__.py:140-156), stop id masked so lengths are fixed, then windows 28/21 -> SNAC."""
    dims = oorph.OrpheusDims.tiny()
    dims.max_tokens = 75
    weights = oorph.synth_weights(dims, seed=3)
    snac_cfg = osnac.SnacConfig.tiny()
    snac_sd = osnac.synth_state_dict(snac_cfg, seed=5)
    cfg = ref.sampling.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1,
                                      greedy=True, max_tokens=dims.max_tokens)
    model = build_ref_orpheus(ref, dims, weights, snac_cfg, snac_sd, cfg)
    page_size, max_pages, max_bs = 16, 64, 4
    worker = build_ref_worker(ref, model, page_size, max_pages, max_bs)

    sched = object.__new__(ref.sched_base.Scheduler)
    sched.model_worker, sched.max_batch_size, sched.active_requests = worker, max_bs, []

    g = torch.Generator().manual_seed(21)
    prompt_lens = [5, 16, 30, 9, 33]
    prompts = [torch.randint(10, dims.vocab_size, (n,), generator=g).tolist() for n in prompt_lens]
    arrivals = {0: [0, 1], 2: [2], 3: [3], 30: [4]}      # step -> request indices joining
    reqs = [ref.requests.Request(request_id=f"r{i}", prompt=p) for i, p in enumerate(prompts)]
    for r, n in zip(reqs, prompt_lens):
        r.input_length = n

    schedule, logits_log, audio = [], [], {r.request_id: [] for r in reqs}
    step = 0
    with _NoisePatch(1234), torch.inference_mode():
        while True:
            for i in arrivals.get(step, []):
                sched.active_requests.append(reqs[i])
            sched.active_requests = [r for r in sched.active_requests if not r.done_all]
            if not sched.active_requests and step > max(arrivals):
                break
            det = sched._select_detokenize_requests()
            lm = sched._select_lm_requests()
            lm_inputs = worker.prepare_lm_inputs(lm, det)
            ref.worker_base.ModelWorker.run_detokenize(worker, det)
            for r in det:
                while not r.output_audio.empty():
                    audio[r.request_id].append(np.frombuffer(r.output_audio.get(), dtype=np.int16))
                if r.done_all:
                    worker.free_kv_cache(r)
            res = ref_lm_step(ref, worker, lm, lm_inputs, stop_mask_id=dims.stop_token_id)
            schedule.append([int(r.request_id[1:]) for r in lm])
            if res is not None and step < 12:
                logits_log.append(res[0][:, 0].float().numpy())
            step += 1
            assert step < 500

    d = {"prompt_lens": np.array(prompt_lens), "n_steps": np.array(step)}
    for i, p in enumerate(prompts):
        d[f"prompt{i}"] = np.array(p)
        d[f"tokens{i}"] = np.array([int(t[0, 0]) for t in reqs[i].lm_output_tokens])
        d[f"audio{i}"] = np.concatenate(audio[f"r{i}"]) if audio[f"r{i}"] else np.zeros(0, np.int16)
        d[f"audio_chunks{i}"] = np.array([len(a) for a in audio[f"r{i}"]])
        d[f"finish{i}"] = np.array(reqs[i].finish_reason or "")
    d["schedule"] = np.array([",".join(map(str, s)) for s in schedule])
    for k, l in enumerate(logits_log):
        d[f"logits_step{k}"] = l
    d["arrival_steps"] = np.array(sorted(arrivals))
    d["arrival_reqs"] = np.array([",".join(map(str, arrivals[s])) for s in sorted(arrivals)])
    np.savez_compressed(os.path.join(OUT, "orpheus_tiny_e2e.npz"), **d)
    print("orpheus_tiny_e2e.npz steps", step, {i: len(d[f'tokens{i}']) for i in range(len(prompts))},
          {i: d[f'audio{i}'].shape for i in range(len(prompts))})
    # margins: how robust is the greedy argmax on these weights?
    tops = [np.sort(l, axis=-1)[:, -2:] for l in logits_log]
    print("min top1-top2 margin over logged steps:", min(float((t[:, 1] - t[:, 0]).min()) for t in tops))


def golden_cosyvoice2_lm():
    """BASELINE.json configs[0] plumbing: the reference's ``CosyVoice2ForCausalLM`` (model/cosyvoice2.py:285-315) at
    a tiny configuration on CPU -- prefill with ``inputs_embeds``, then greedy decode steps fed with
    ``speech_embedding(id)`` -- through the paged CPU wrapper.  Stored: the prompt embeddings' seed, per-step logits
    (fp32 copies of the bf16 outputs) and the greedy ids."""
    from . import cosyvoice2 as ocv
    from .ref_import import import_reference_cosyvoice2

    mod = import_reference_cosyvoice2()
    dims = ocv.CosyVoice2Dims.tiny()
    cfg = mod.CosyVoice2Config(hidden_size=dims.hidden_size, intermediate_size=dims.intermediate_size,
                               num_attention_heads=dims.num_attention_heads,
                               num_key_value_heads=dims.num_key_value_heads,
                               num_hidden_layers=dims.num_hidden_layers, vocab_size=dims.vocab_size,
                               rope_theta=dims.rope_theta, rms_norm_eps=dims.rms_norm_eps)
    cfg.llm_input_size = cfg.llm_output_size = dims.hidden_size
    cfg.speech_token_size = dims.speech_token_size
    weights = ocv.synth_weights(dims, seed=3)
    lm = mod.CosyVoice2ForCausalLM(cfg)
    lm.load_state_dict(weights, strict=True)
    lm = lm.to(torch.bfloat16).eval()
    page_size, T0, n_steps, seed = 16, 21, 14, 1        # 21 + 14 tokens cross two page boundaries
    emb = torch.randn(T0, dims.hidden_size, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16)
    n_pages = (T0 + n_steps + page_size - 1) // page_size + 1
    kv = torch.zeros(dims.num_hidden_layers, n_pages, 2, page_size, dims.num_key_value_heads, dims.head_dim,
                     dtype=torch.bfloat16)
    pages = list(range((T0 + page_size - 1) // page_size))
    with torch.no_grad():
        pre = lm_ops.PagedWrapperCPU("prefill", page_size)
        pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
        logits = lm(inputs_embeds=emb, position_ids=torch.arange(T0, dtype=torch.int32), attn_wrapper=pre,
                    kv_cache=kv)[-1:]
        ids, logs, kv_len = [], [logits[0].float().numpy()], T0
        for _ in range(n_steps):
            tok = int(torch.argmax(logits[0].float()))
            ids.append(tok)
            kv_len += 1
            if (kv_len + page_size - 1) // page_size > len(pages):
                pages.append(len(pages))
            dec = lm_ops.PagedWrapperCPU("decode", page_size)
            dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
            step_emb = lm.embed_tokens_speech(torch.tensor([tok]))
            logits = lm(inputs_embeds=step_emb, position_ids=torch.tensor([kv_len - 1], dtype=torch.int32),
                        attn_wrapper=dec, kv_cache=kv)
            logs.append(logits[0].float().numpy())
    np.savez_compressed(os.path.join(OUT, "cosyvoice2_tiny_lm.npz"), ids=np.array(ids, dtype=np.int64),
                        logits=np.stack(logs), prompt_seed=seed, prompt_len=T0, page_size=page_size, weight_seed=3)
    tops = np.sort(np.stack(logs), axis=-1)[:, -2:]
    print("cosyvoice2_tiny_lm.npz ids", ids, "min top1-top2 margin", float((tops[:, 1] - tops[:, 0]).min()))


def golden_glm_voice_lm():
    """BASELINE.json configs[4]'s decoder: the reference's ``GLMVoiceForCausalLM`` (model/glm_voice.py:281-304) at a
    tiny configuration on CPU -- fused biased QKV, interleaved partial RoPE, fused SwiGLU -- prefill on prompt ids,
    then greedy decode steps through the paged CPU wrapper."""
    import importlib

    from . import glm_voice as oglm

    import_reference()
    mod = importlib.import_module("vox_serve.model.glm_voice")
    mod.rms_norm = lambda hidden_states, weight, eps: lm_ops.rms_norm(hidden_states, weight, eps)
    mod.apply_rope_pos_ids = (
        lambda query_states, key_states, position_ids, **kw: lm_ops.apply_rope_pos_ids(
            query_states, key_states, position_ids, **kw))
    dims = oglm.GLMVoiceDims.tiny()
    cfg = mod.GLMVoiceConfig(ffn_hidden_size=dims.ffn_hidden_size, hidden_size=dims.hidden_size,
                             layernorm_epsilon=dims.layernorm_epsilon,
                             multi_query_group_num=dims.multi_query_group_num,
                             num_attention_heads=dims.num_attention_heads, num_hidden_layers=dims.num_layers,
                             num_layers=dims.num_layers, padded_vocab_size=dims.padded_vocab_size,
                             vocab_size=dims.padded_vocab_size, rope_ratio=int(dims.rope_ratio))
    weights = oglm.synth_weights(dims, seed=4)
    lm = mod.GLMVoiceForCausalLM(cfg)
    lm.load_state_dict(weights, strict=True)
    lm = lm.to(torch.bfloat16).eval()
    page_size, T0, n_steps, seed = 16, 27, 12, 2           # 27 + 12 tokens cross two page boundaries
    prompt = torch.randint(0, dims.padded_vocab_size, (T0,), generator=torch.Generator().manual_seed(seed))
    n_pages = (T0 + n_steps + page_size - 1) // page_size + 1
    kv = torch.zeros(dims.num_layers, n_pages, 2, page_size, dims.multi_query_group_num, dims.head_dim,
                     dtype=torch.bfloat16)
    pages = list(range((T0 + page_size - 1) // page_size))
    with torch.no_grad():
        pre = lm_ops.PagedWrapperCPU("prefill", page_size)
        pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
        logits = lm(inputs_embeds=lm.embed_tokens(prompt), position_ids=torch.arange(T0, dtype=torch.int32),
                    attn_wrapper=pre, kv_cache=kv)[-1:]
        ids, logs, kv_len = [], [logits[0].float().numpy()], T0
        for _ in range(n_steps):
            tok = int(torch.argmax(logits[0].float()))
            ids.append(tok)
            kv_len += 1
            if (kv_len + page_size - 1) // page_size > len(pages):
                pages.append(len(pages))
            dec = lm_ops.PagedWrapperCPU("decode", page_size)
            dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
            logits = lm(inputs_embeds=lm.embed_tokens(torch.tensor([tok])),
                        position_ids=torch.tensor([kv_len - 1], dtype=torch.int32), attn_wrapper=dec, kv_cache=kv)
            logs.append(logits[0].float().numpy())
    np.savez_compressed(os.path.join(OUT, "glm_voice_tiny_lm.npz"), ids=np.array(ids, dtype=np.int64),
                        logits=np.stack(logs), prompt=prompt.numpy(), page_size=page_size, weight_seed=4)
    tops = np.sort(np.stack(logs), axis=-1)[:, -2:]
    print("glm_voice_tiny_lm.npz ids", ids, "min top1-top2 margin", float((tops[:, 1] - tops[:, 0]).min()))


def golden_csm_frames():
    """BASELINE.json configs[3] / SURVEY row a24: the reference's ``CsmBackboneModel``, ``CsmDepthDecoderForCausalLM``
    and ``CsmCodebooksHead`` (model/csm.py:171-272) at a tiny configuration on CPU, driven the way
    ``CSMModel.forward / sampling / depth_sampling`` (:637-769) and ``CudaGraphWorker.run_lm_depth``
    (cuda_graph_worker.py:1058-1160) drive them: backbone prefill on a text + audio prompt, then greedy frames --
    codebook 0 from the backbone, the 2-row depth prefill and the 1-row depth decodes on a zeroed per-frame cache."""
    import importlib

    from torch import nn
    from transformers import CsmConfig

    from . import csm as ocsm

    import_reference()
    mod = importlib.import_module("vox_serve.model.csm")
    mod.rms_norm = lambda hidden_states, weight, eps: lm_ops.rms_norm(hidden_states, weight, eps)
    mod.apply_rope_pos_ids = (
        lambda query_states, key_states, position_ids, **kw: lm_ops.apply_rope_pos_ids(
            query_states, key_states, position_ids, **kw))
    d = ocsm.CsmDims.tiny()
    dcfg = dict(num_codebooks=d.num_codebooks, vocab_size=d.vocab_size, backbone_hidden_size=d.hidden_size,
                hidden_size=d.depth_hidden_size, intermediate_size=d.depth_intermediate_size,
                num_hidden_layers=d.depth_num_hidden_layers, num_attention_heads=d.depth_num_attention_heads,
                num_key_value_heads=d.depth_num_key_value_heads, head_dim=d.depth_head_dim, rms_norm_eps=d.rms_norm_eps)
    cfg = CsmConfig(num_codebooks=d.num_codebooks, vocab_size=d.vocab_size, text_vocab_size=d.text_vocab_size,
                    hidden_size=d.hidden_size, intermediate_size=d.intermediate_size,
                    num_hidden_layers=d.num_hidden_layers, num_attention_heads=d.num_attention_heads,
                    num_key_value_heads=d.num_key_value_heads, head_dim=d.head_dim, rms_norm_eps=d.rms_norm_eps,
                    depth_decoder_config=dcfg)
    for c in (cfg, cfg.depth_decoder_config):       # transformers 5.x keeps theta inside rope_parameters
        c.rope_theta = d.rope_theta
    w = ocsm.synth_weights(d, seed=6)
    backbone = mod.CsmBackboneModel(cfg)
    depth = mod.CsmDepthDecoderForCausalLM(cfg.depth_decoder_config)
    lm_head = nn.Linear(d.hidden_size, d.vocab_size, bias=False)
    text_emb = nn.Embedding(d.text_vocab_size, d.hidden_size)
    backbone.load_state_dict({k[len("backbone_model."):]: v for k, v in w.items() if k.startswith("backbone_model.")},
                             strict=True)
    depth.load_state_dict({k[len("depth_decoder."):]: v for k, v in w.items() if k.startswith("depth_decoder.")},
                          strict=True)
    lm_head.load_state_dict({"weight": w["lm_head.weight"]})
    text_emb.load_state_dict({"weight": w["embed_text_tokens.weight"]})
    for m_ in (backbone, depth, lm_head, text_emb):
        m_.to(torch.bfloat16).eval()
    N, V = d.num_codebooks, d.vocab_size
    g = torch.Generator().manual_seed(3)
    T0, n_text, n_frames, page_size, depth_page = 19, 12, 5, 16, 32
    ids = torch.zeros(T0, N + 1, dtype=torch.long)
    masks = torch.zeros(T0, N + 1, dtype=torch.bool)
    ids[:n_text, -1] = torch.randint(0, d.text_vocab_size, (n_text,), generator=g)
    masks[:n_text, -1] = True
    ids[n_text:, :-1] = torch.randint(0, V, (T0 - n_text, N), generator=g)
    masks[n_text:, :-1] = True

    def embeds(row_ids, row_masks):          # CSMModel.forward, csm.py:647-654
        e = torch.cat([backbone.embed_tokens(row_ids[:, :-1]), text_emb(row_ids[:, -1:])], dim=1)
        return (e * row_masks[:, :, None]).sum(dim=1)

    def single(tok, i):                      # embed_audio_tokens_single, csm.py:456-458
        return backbone.embed_tokens.embed_audio_tokens(tok + i * V)

    n_pages = (T0 + n_frames + page_size - 1) // page_size + 1
    kv = torch.zeros(d.num_hidden_layers, n_pages, 2, page_size, d.num_key_value_heads, d.head_dim, dtype=torch.bfloat16)
    dkv = torch.zeros(d.depth_num_hidden_layers, 1, 2, depth_page, d.depth_num_key_value_heads, d.depth_head_dim,
                      dtype=torch.bfloat16)
    pages = list(range((T0 + page_size - 1) // page_size))
    frames, cb0_logits, depth_logits = [], [], []
    with torch.no_grad():
        pre = lm_ops.PagedWrapperCPU("prefill", page_size)
        pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
        hidden = backbone(embeds(ids, masks), torch.arange(T0, dtype=torch.int32), pre, kv)
        logits, hidden = lm_head(hidden)[-1], hidden[-1]
        kv_len = T0
        dmask = torch.ones(1, N + 1, dtype=torch.bool)
        dmask[0, -1] = False
        for _ in range(n_frames):
            cb0_logits.append(logits.float().numpy())
            cb0 = torch.argmax(logits.float()).view(1)
            frame, dl = [int(cb0)], []
            dkv.zero_()                                                                   # :1076
            x = torch.cat([hidden[None, None, :], single(cb0, 0)[:, None, :]], dim=1).view(2, -1)   # :701-702
            dpre = lm_ops.PagedWrapperCPU("prefill", depth_page)
            dpre.plan([0, 2], [0, 1], [0], [2])
            pos = torch.tensor([0, 1], dtype=torch.int32)
            out = depth.codebooks_head(depth(x, pos, dpre, dkv), cache_position=pos)[-1:]
            for i in range(1, N):
                dl.append(out[0].float().numpy())
                tok = torch.argmax(out[0].float()).view(1)
                frame.append(int(tok))
                if i == N - 1:
                    break
                ddec = lm_ops.PagedWrapperCPU("decode", depth_page)
                ddec.plan([0, 1], [0], [i + 2])
                pos = torch.tensor([i + 1], dtype=torch.int32)
                out = depth.codebooks_head(depth(single(tok, i), pos, ddec, dkv), cache_position=pos)
            frames.append(frame)
            depth_logits.append(np.stack(dl))
            kv_len += 1
            if (kv_len + page_size - 1) // page_size > len(pages):
                pages.append(len(pages))
            dec = lm_ops.PagedWrapperCPU("decode", page_size)
            dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
            row = torch.tensor([frame + [0]], dtype=torch.long)
            hidden = backbone(embeds(row, dmask), torch.tensor([kv_len - 1], dtype=torch.int32), dec, kv)
            logits, hidden = lm_head(hidden)[0], hidden[0]
    np.savez_compressed(os.path.join(OUT, "csm_tiny_frames.npz"), frames=np.array(frames, dtype=np.int64),
                        cb0_logits=np.stack(cb0_logits), depth_logits=np.stack(depth_logits), prompt_ids=ids.numpy(),
                        prompt_masks=masks.numpy(), page_size=page_size, weight_seed=6)
    print("csm_tiny_frames.npz frames", frames)


def golden_qwen3_tts_frames():
    """BASELINE.json configs[2] / SURVEY row a24: the reference's ``Qwen3TTSTalkerForConditionalGeneration`` (talker,
    text projection, codec head, code predictor; model/qwen3_tts.py:707-833) at a tiny configuration on CPU, driven
    through ``Qwen3TTSForCausalLM.forward / forward_depth`` (:906-944, called unbound on a namespace holding the
    talker) with the glue of ``Qwen3TTSModel.forward / sampling / depth_sampling`` (:1805-2004) and the worker's
    depth loop (cuda_graph_worker.py:1058-1160)."""
    from . import qwen3_tts as oq
    from .ref_import import import_reference_qwen3_tts

    mod = import_reference_qwen3_tts()
    d = oq.Qwen3TTSDims.tiny()
    cp = mod.Qwen3TTSCodePredictorConfig(head_dim=d.cp_head_dim, hidden_size=d.cp_hidden_size,
                                         intermediate_size=d.cp_intermediate_size,
                                         num_attention_heads=d.cp_num_attention_heads,
                                         num_code_groups=d.num_code_groups, num_hidden_layers=d.cp_num_hidden_layers,
                                         num_key_value_heads=d.cp_num_key_value_heads, rms_norm_eps=d.rms_norm_eps,
                                         rope_theta=int(d.rope_theta), vocab_size=d.cp_vocab_size)
    tk = mod.Qwen3TTSTalkerConfig(code_predictor_config=cp, head_dim=d.head_dim, hidden_size=d.hidden_size,
                                  intermediate_size=d.intermediate_size, num_attention_heads=d.num_attention_heads,
                                  num_code_groups=d.num_code_groups, num_hidden_layers=d.num_hidden_layers,
                                  num_key_value_heads=d.num_key_value_heads, rms_norm_eps=d.rms_norm_eps,
                                  rope_theta=int(d.rope_theta), text_hidden_size=d.text_hidden_size,
                                  text_vocab_size=d.text_vocab_size, vocab_size=d.vocab_size)
    w = oq.synth_weights(d, seed=8)
    talker = mod.Qwen3TTSTalkerForConditionalGeneration(tk)
    r = talker.load_state_dict({k[len("talker."):]: v for k, v in w.items()}, strict=False)
    assert not r.unexpected_keys and r.missing_keys == ["code_predictor.lm_head_weight"], (r.missing_keys, r.unexpected_keys)
    talker = talker.to(torch.bfloat16).eval()
    talker.code_predictor.lm_head_weight = torch.stack([h.weight.detach() for h in talker.code_predictor.lm_head], 0)
    lm = types.SimpleNamespace(talker=talker)                        # what Qwen3TTSForCausalLM.forward* use of self
    fwd, fwd_depth = mod.Qwen3TTSForCausalLM.forward, mod.Qwen3TTSForCausalLM.forward_depth
    N = d.num_code_groups
    g = torch.Generator().manual_seed(5)
    T0, n_frames, page_size, depth_page = 18, 5, 16, 32
    text = torch.randint(0, d.text_vocab_size, (T0,), generator=g)
    cb0p = torch.randint(0, d.vocab_size, (T0,), generator=g)
    need = torch.zeros(T0, dtype=torch.bool)
    need[10:] = True
    feat = torch.zeros(T0, d.hidden_size, dtype=torch.bfloat16)
    feat[3] = torch.randn(d.hidden_size, generator=g).to(torch.bfloat16)      # a speaker-embedding position

    def embeds(text_ids, cb0_ids, needs_codec, feats):                # Qwen3TTSModel.forward :1835-1853
        t = talker.text_projection(talker.model.text_embedding(text_ids))
        c = talker.model.codec_embedding(cb0_ids)
        return torch.where(needs_codec.unsqueeze(-1), t + c, t) + feats

    n_pages = (T0 + n_frames + page_size - 1) // page_size + 1
    kv = torch.zeros(d.num_hidden_layers, n_pages, 2, page_size, d.num_key_value_heads, d.head_dim, dtype=torch.bfloat16)
    dkv = torch.zeros(d.cp_num_hidden_layers, 1, 2, depth_page, d.cp_num_key_value_heads, d.cp_head_dim,
                      dtype=torch.bfloat16)
    pages = list(range((T0 + page_size - 1) // page_size))
    frames, cb0_logits, cp_logits = [], [], []
    with torch.no_grad():
        pre = lm_ops.PagedWrapperCPU("prefill", page_size)
        pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
        logits, hidden = fwd(lm, embeds(text, cb0p, need, feat), torch.arange(T0, dtype=torch.int32), pre, kv)
        logits, hidden = logits[-1], hidden[-1]
        kv_len = T0
        for _ in range(n_frames):
            cb0_logits.append(logits.float().numpy())
            cb0 = torch.argmax(logits.float()).view(1)
            frame, cl = [int(cb0)], []
            dkv.zero_()
            x = torch.cat([hidden[None, None, :], talker.model.codec_embedding(cb0)[:, None, :]], dim=1).view(2, -1)
            dpre = lm_ops.PagedWrapperCPU("prefill", depth_page)
            dpre.plan([0, 2], [0, 1], [0], [2])
            out = fwd_depth(lm, x, torch.tensor([0, 1], dtype=torch.int32), dpre, dkv)[-1:]
            in_feat = torch.zeros(1, d.hidden_size, dtype=torch.bfloat16)          # :1942
            for i in range(1, N):
                cl.append(out[0].float().numpy())
                tok = torch.argmax(out[0].float()).view(1)
                frame.append(int(tok))
                ci = talker.code_predictor.model.codec_embedding[i - 1](tok)       # depth_sampling :1995
                in_feat[:] += ci                                                   # :2002
                if i == N - 1:
                    break
                ddec = lm_ops.PagedWrapperCPU("decode", depth_page)
                ddec.plan([0, 1], [0], [i + 2])
                out = fwd_depth(lm, ci, torch.tensor([i + 1], dtype=torch.int32), ddec, dkv)
            frames.append(frame)
            cp_logits.append(np.stack(cl))
            kv_len += 1
            if (kv_len + page_size - 1) // page_size > len(pages):
                pages.append(len(pages))
            dec = lm_ops.PagedWrapperCPU("decode", page_size)
            dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
            e = embeds(torch.tensor([d.tts_pad_token_id]), cb0, torch.tensor([True]), in_feat)
            logits, hidden = fwd(lm, e, torch.tensor([kv_len - 1], dtype=torch.int32), dec, kv)
            logits, hidden = logits[0], hidden[0]
    np.savez_compressed(os.path.join(OUT, "qwen3_tts_tiny_frames.npz"), frames=np.array(frames, dtype=np.int64),
                        cb0_logits=np.stack(cb0_logits), cp_logits=np.stack(cp_logits), text=text.numpy(),
                        cb0=cb0p.numpy(), needs_codec=need.numpy(), features=feat.float().numpy(),
                        page_size=page_size, weight_seed=8)
    print("qwen3_tts_tiny_frames.npz frames", frames)


def golden_mimi():
    """The reference's own ``MimiModel.decode`` (tokenizer/mimi.py:2993-3018) on CPU at a tiny configuration with all four
    SEANet ratios, seeded weights from oracle.mimi.synth_state_dict loaded with ``load_state_dict(strict=False)`` (the
    encoder side keeps its default initialisation: decode never touches it) -> tests/golden/mimi_tiny.npz."""
    from . import mimi as omimi
    from .ref_import import REFERENCE_ROOT

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from vox_serve.tokenizer import mimi as rm

    cfg = omimi.MimiConfig.tiny()
    seanet = dict(rm._seanet_kwargs)
    seanet.update(dimension=cfg.dimension, n_filters=cfg.n_filters, ratios=list(cfg.ratios))
    quant = dict(dimension=cfg.codebook_dim, n_q=cfg.n_q, bins=cfg.bins, input_dimension=cfg.dimension,
                 output_dimension=cfg.dimension)
    tr = dict(rm._transformer_kwargs)
    tr.update(d_model=cfg.dimension, num_heads=cfg.num_heads, num_layers=cfg.num_layers, dim_feedforward=cfg.dim_feedforward,
              input_dimension=cfg.dimension, output_dimensions=[cfg.dimension])
    enc, dec = rm.SEANetEncoder(**seanet), rm.SEANetDecoder(**seanet)
    model = rm.MimiModel(enc, dec, rm.SplitResidualVectorQuantizer(**quant), channels=1, sample_rate=24000, frame_rate=12.5,
                         encoder_frame_rate=24000 / enc.hop_length, causal=True, resample_method="conv",
                         encoder_transformer=rm.ProjectedTransformer(device="cpu", **tr),
                         decoder_transformer=rm.ProjectedTransformer(device="cpu", **tr)).eval()
    seed = 17
    sd = omimi.synth_state_dict(cfg, seed)
    r = model.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all(k.startswith(("encoder", "downsample")) for k in r.missing_keys), r.missing_keys
    model.set_num_codebooks(cfg.n_q)
    g = torch.Generator().manual_seed(3)
    codes = torch.randint(0, cfg.bins, (3, cfg.n_q, 5), generator=g)
    with torch.no_grad():
        wav = model.decode(codes)
        latent = model._to_encoder_framerate(model.decode_latent(codes))
        (tr_out,) = model.decoder_transformer(latent)
    assert wav.shape == (3, 1, 5 * cfg.hop)
    np.savez_compressed(os.path.join(OUT, "mimi_tiny.npz"), weight_seed=seed, codes=codes.numpy(), wav=wav.numpy(),
                        latent=latent.numpy(), transformer_out=tr_out.numpy())
    print("mimi_tiny.npz:", tuple(wav.shape), "abs max %.4f" % float(wav.abs().max()))


def golden_qwen3_codec():
    """The reference's own ``Qwen3TTSTokenizerV2Decoder`` (tokenizer/qwen3_codec.py:1307-1667) on CPU at a tiny configuration
    (GQA 2, sliding window 12, all four decoder rates): ``init_cache`` + three consecutive ``forward_chunk`` calls of 5, 5 and
    3 frames -- the window wraps, the short last chunk takes the "chunk shorter than the conv cache" branch of the
    dilation-9 convolutions -> tests/golden/qwen3_codec_tiny.npz (waveforms + the final cache)."""
    from . import qwen3_codec as oq
    from .ref_import import REFERENCE_ROOT

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from vox_serve.tokenizer import qwen3_codec as rq

    rq.ROPE_INIT_FUNCTIONS = None      # transformers 5 dropped the "default" entry; the module's own fallback is the same formula
    cfg = oq.Qwen3CodecConfig.tiny()
    rcfg = rq.Qwen3TTSTokenizerV2DecoderConfig(
        latent_dim=cfg.latent_dim, codebook_dim=cfg.codebook_dim, codebook_size=cfg.codebook_size, decoder_dim=cfg.decoder_dim,
        hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, head_dim=cfg.head_dim,
        num_attention_heads=cfg.num_attention_heads, num_hidden_layers=cfg.num_hidden_layers,
        num_key_value_heads=cfg.num_key_value_heads, num_quantizers=cfg.num_quantizers, rms_norm_eps=cfg.rms_norm_eps,
        rope_theta=cfg.rope_theta, sliding_window=cfg.sliding_window, upsample_rates=list(cfg.upsample_rates),
        upsampling_ratios=list(cfg.upsampling_ratios))
    dec = rq.Qwen3TTSTokenizerV2Decoder(rcfg).eval()
    seed, B = 23, 2
    sd = oq.synth_state_dict(cfg, seed)
    dec.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(5)
    chunks = [torch.randint(0, cfg.codebook_size, (B, cfg.num_quantizers, t), generator=g) for t in (5, 5, 3)]
    out = {}
    with torch.no_grad():
        cache = dec.init_cache(B, torch.device("cpu"), torch.float32, detokenize_interval=5)
        for i, c in enumerate(chunks):
            if c.shape[2] != 5:       # the work buffers are sized per chunk length: the state tensors carry over
                fresh = dec.init_cache(B, torch.device("cpu"), torch.float32, detokenize_interval=c.shape[2])
                for name in ("attention_cache", "position_offset", "pre_conv_cache"):
                    getattr(fresh, name).copy_(getattr(cache, name))
                for name in ("upsample_conv_caches", "decoder_conv_caches", "transconv_caches"):
                    for a, b in zip(getattr(fresh, name), getattr(cache, name)):
                        a.copy_(b)
                cache = fresh
            wav, cache = dec.forward_chunk(c, cache)
            assert wav.shape == (B, 1, c.shape[2] * cfg.hop), wav.shape
            out[f"codes{i}"], out[f"wav{i}"] = c.numpy(), wav.numpy().copy()
    out["attention_cache"], out["position_offset"] = cache.attention_cache.numpy(), cache.position_offset.numpy()
    out["pre_conv_cache"] = cache.pre_conv_cache.numpy()
    for name in ("upsample_conv_caches", "decoder_conv_caches", "transconv_caches"):
        for j, t in enumerate(getattr(cache, name)):
            out[f"{name}.{j}"] = t.numpy()
    np.savez_compressed(os.path.join(OUT, "qwen3_codec_tiny.npz"), weight_seed=seed, **out)
    print("qwen3_codec_tiny.npz:", [tuple(out[f"wav{i}"].shape) for i in range(3)],
          "abs max %.4f" % max(float(np.abs(out[f"wav{i}"]).max()) for i in range(3)))


def golden_glm_encoder():
    """The reference's own ``GLMWhisperVQEncoder`` (vox_serve/encoder/glm.py:217-323) in bf16 on CPU, tiny config,
    seeded weights from oracle.glm_encoder.synth_state_dict; two inputs: a full-length one and one whose tail is
    padding (attention_mask 0) with a length that is not a multiple of the pooling width after the convolutions."""
    from .ref_import import REFERENCE_ROOT

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from vox_serve.encoder.glm import GLMEncoderConfig, GLMWhisperVQEncoder

    from . import glm_encoder as oenc

    d = oenc.GLMEncoderDims.tiny()
    seed = 11
    sd = oenc.synth_state_dict(d, seed)
    cfg = GLMEncoderConfig(d_model=d.d_model, encoder_attention_heads=d.encoder_attention_heads,
                           encoder_ffn_dim=d.encoder_ffn_dim, num_mel_bins=d.num_mel_bins,
                           max_source_positions=d.max_source_positions, pooling_kernel_size=d.pooling_kernel_size,
                           pooling_position=d.pooling_position, quantize_position=d.quantize_position,
                           quantize_vocab_size=d.quantize_vocab_size,
                           quantize_causal_block_size=d.quantize_causal_block_size)
    m = GLMWhisperVQEncoder(cfg).to(torch.bfloat16).eval()
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    out = {"weight_seed": seed}
    g = torch.Generator().manual_seed(5)
    cases = {"full": (160, 160), "padded": (148, 132)}          # (frames, valid frames)
    for tag, (frames, valid) in cases.items():
        feats = torch.randn(1, d.num_mel_bins, frames, generator=g).to(torch.bfloat16)
        mask = torch.zeros(1, frames, dtype=torch.long)
        mask[:, :valid] = 1
        states = {}
        hook = m.layers[-1].register_forward_hook(lambda mod, i, o: states.__setitem__("last", o))
        with torch.no_grad():
            ids = m(feats, mask)
        hook.remove()
        out[f"{tag}_features"] = feats.float().numpy()
        out[f"{tag}_mask"] = mask.numpy()
        out[f"{tag}_ids"] = ids.numpy()
        out[f"{tag}_hidden"] = states["last"].float().numpy()
    np.savez_compressed(os.path.join(OUT, "glm_encoder_tiny.npz"), **out)
    print("glm_encoder_tiny.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape")},
          "distinct ids:", len(set(out["full_ids"].reshape(-1).tolist())))


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "glm_encoder":
        golden_glm_encoder()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "qwen3_codec":
        golden_qwen3_codec()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "qwen3_tts":
        golden_qwen3_tts_frames()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "csm":
        golden_csm_frames()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "mimi":
        golden_mimi()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "cosyvoice2":
        golden_cosyvoice2_lm()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "glm_voice":
        golden_glm_voice_lm()
        return
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = import_reference()
    orig_sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        golden_sampler(ref)
        golden_snac(ref)
        golden_orpheus_e2e(ref)
        golden_cosyvoice2_lm()
        golden_glm_voice_lm()
        golden_csm_frames()
        golden_qwen3_tts_frames()
        golden_mimi()
        golden_qwen3_codec()
        golden_glm_encoder()
    finally:
        torch.cuda.synchronize = orig_sync


if __name__ == "__main__":
    main()
