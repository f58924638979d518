"""The CPU oracle (oracle/*.py restatements) held to the golden vectors that
oracle/gen_golden.py produced by running the reference's own code (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import orpheus as oorph, sampler as osampler, snac as osnac, worker as oworker


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_sampler_penalty_and_greedy(golden_dir):
    g = _load(golden_dir, "sampler.npz")
    logits = torch.from_numpy(g["pen_logits"]).to(torch.bfloat16)
    cache = torch.from_numpy(g["pen_cache"])
    out = osampler.apply_repetition_penalty(logits, cache, 1.3)
    assert np.array_equal(out.float().numpy(), g["pen_out"])
    ids = osampler.greedy(out.view(-1, out.shape[-1]))
    assert np.array_equal(ids.numpy(), g["greedy_ids"])


def test_sampler_cache_update_global_marks_batch_union(golden_dir):
    g = _load(golden_dir, "sampler.npz")
    cache = torch.from_numpy(g["pen_cache"]).clone()
    ids = torch.from_numpy(g["upd_global_ids"])
    osampler.update_repetition_cache(cache, ids, -1)
    assert np.array_equal(cache.numpy(), g["upd_global_out"])
    # the quirk itself: every row carries every row's id
    for b in range(cache.shape[0]):
        assert cache[b, 0, 0, ids[:, 0]].all()


def test_sampler_windowed_multicodebook(golden_dir):
    g = _load(golden_dir, "sampler.npz")
    cw = torch.from_numpy(g["upd_win_in"])
    idw = torch.from_numpy(g["upd_win_ids"])
    c = cw.clone()
    osampler.update_repetition_cache(c, idw, 3)
    assert np.array_equal(c.numpy(), g["upd_win_out"])
    lw = torch.from_numpy(g["pen_win_logits"]).to(torch.bfloat16)
    assert np.array_equal(osampler.apply_repetition_penalty(lw, cw, 1.7).float().numpy(), g["pen_win_out"])
    assert np.array_equal(osampler.apply_repetition_penalty(lw[:, :1], cw, 1.7).float().numpy(), g["pen_cb0_out"])
    c = cw.clone()
    osampler.update_repetition_cache(c, idw[:, :1], 3)
    assert np.array_equal(c.numpy(), g["upd_cb0_win_out"])
    c = cw.clone()
    osampler.update_repetition_cache(c, idw[:, :1], -1)
    assert np.array_equal(c.numpy(), g["upd_cb0_global_out"])


@pytest.mark.parametrize("tag,cfg", [("tiny", osnac.SnacConfig.tiny()), ("24khz", osnac.SnacConfig())])
def test_snac_decode(golden_dir, tag, cfg):
    g = _load(golden_dir, f"snac_{tag}.npz")
    sd = osnac.synth_state_dict(cfg, seed=5)
    codes = [torch.from_numpy(g[f"codes{i}"]) for i in range(3)]
    noises = [torch.from_numpy(g[f"noise{i}"]) for i in range(4)]
    wav = osnac.decode(sd, cfg, codes, noises).numpy()
    assert wav.shape == g["wav"].shape
    np.testing.assert_allclose(wav, g["wav"], atol=2e-5, rtol=0)


def test_orpheus_tiny_end_to_end(golden_dir):
    g = _load(golden_dir, "orpheus_tiny_e2e.npz")
    dims = oorph.OrpheusDims.tiny()
    dims.max_tokens = 75
    weights = oorph.synth_weights(dims, seed=3)
    scfg = osnac.SnacConfig.tiny()
    ssd = osnac.synth_state_dict(scfg, seed=5)
    cfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1,
                                  greedy=True, max_tokens=75)
    w = oworker.OracleWorker(weights, dims, cfg, page_size=16, max_num_pages=64, snac_sd=ssd, snac_cfg=scfg,
                             max_batch_size=4, noise_seed=1234, ignore_stop=True)
    n_req = len(g["prompt_lens"])
    reqs = [oworker.Req(f"r{i}", torch.from_numpy(g[f"prompt{i}"])) for i in range(n_req)]
    arrivals = {int(s): [int(x) for x in str(r).split(",")] for s, r in zip(g["arrival_steps"], g["arrival_reqs"])}
    active, schedule = [], []
    for step in range(int(g["n_steps"])):
        for i in arrivals.get(step, []):
            active.append(reqs[i])
        active = [r for r in active if not r.done_all]
        lm, _ = w.step(active)
        schedule.append(",".join(str(int(r.request_id[1:])) for r in lm))
        if f"logits_step{step}" in g.files:
            assert np.array_equal(w.last_logits[:, 0].float().numpy(), g[f"logits_step{step}"]), step
    assert schedule == list(g["schedule"])
    for i, r in enumerate(reqs):
        assert r.lm_output_tokens == g[f"tokens{i}"].tolist(), i
        assert r.finish_reason == str(g[f"finish{i}"])
        audio = np.concatenate([np.frombuffer(b, dtype=np.int16) for b in r.output_audio])
        assert [len(b) // 2 for b in r.output_audio] == g[f"audio_chunks{i}"].tolist()
        assert audio.shape == g[f"audio{i}"].shape
        diff = np.abs(audio.astype(np.int32) - g[f"audio{i}"].astype(np.int32))
        assert diff.max() <= 1, (i, diff.max())


def test_cosyvoice2_lm_oracle_matches_reference_golden(golden_dir):
    """BASELINE.json configs[0] (CosyVoice2 speech LM, single prompt, greedy, CPU): oracle/cosyvoice2.py against the
    logits and ids the reference's own CosyVoice2ForCausalLM produced (oracle/gen_golden.py:golden_cosyvoice2_lm) --
    q/k/v bias, plain RoPE at theta 1e6, inputs_embeds prefill, speech-embedding feedback, biased llm_decoder, paged
    KV growing across two page boundaries.  Same torch CPU kernels on both sides: bit-exact."""
    from oracle import cosyvoice2 as ocv

    gd = _load(golden_dir, "cosyvoice2_tiny_lm.npz")
    dims = ocv.CosyVoice2Dims.tiny()
    w = ocv.synth_weights(dims, seed=int(gd["weight_seed"]))
    emb = torch.randn(int(gd["prompt_len"]), dims.hidden_size,
                      generator=torch.Generator().manual_seed(int(gd["prompt_seed"]))).to(torch.bfloat16)
    out = ocv.greedy_decode(w, dims, emb, len(gd["ids"]), page_size=int(gd["page_size"]), stop_ids=[])
    assert out["ids"] == gd["ids"].tolist()
    got = torch.stack(out["logits"]).numpy()
    assert got.shape == gd["logits"].shape
    assert np.array_equal(got, gd["logits"])
    # the full-size dims are the reference's (cosyvoice2.py:26-38)
    full = ocv.CosyVoice2Dims()
    assert (full.hidden_size, full.num_hidden_layers, full.num_attention_heads, full.num_key_value_heads,
            full.head_dim, full.intermediate_size, full.speech_vocab) == (896, 24, 14, 2, 64, 4864, 6564)


def test_glm_voice_lm_oracle_matches_reference_golden(golden_dir):
    """BASELINE.json configs[4]'s decoder (GLM-4-Voice): oracle/glm_voice.py against the logits and ids of the
    reference's own GLMVoiceForCausalLM on CPU (oracle/gen_golden.py:golden_glm_voice_lm) -- fused biased QKV,
    interleaved RoPE on half of head_dim, fused SwiGLU, paged KV growing across two page boundaries.  Bit-exact."""
    from oracle import glm_voice as oglm

    gd = _load(golden_dir, "glm_voice_tiny_lm.npz")
    dims = oglm.GLMVoiceDims.tiny()
    w = oglm.synth_weights(dims, seed=int(gd["weight_seed"]))
    out = oglm.greedy_decode(w, dims, torch.from_numpy(gd["prompt"]), len(gd["ids"]), page_size=int(gd["page_size"]))
    assert out["ids"] == gd["ids"].tolist()
    assert np.array_equal(torch.stack(out["logits"]).numpy(), gd["logits"])
    full = oglm.GLMVoiceDims()       # the reference's dims (glm_voice.py:22-54)
    assert (full.hidden_size, full.num_layers, full.num_attention_heads, full.multi_query_group_num, full.head_dim,
            full.ffn_hidden_size, full.padded_vocab_size) == (4096, 40, 32, 2, 128, 13696, 168960)


def test_csm_depth_loop_oracle_matches_reference_golden(golden_dir):
    """BASELINE.json configs[3] / SURVEY row a24 (the first "next" row): oracle/csm.py -- masked multi-codebook frame
    embedding, Llama-style backbone, 2-row depth prefill + 1-row depth decodes with the per-position codebook heads on
    a per-frame cache -- against the reference's own CsmBackboneModel / CsmDepthDecoderForCausalLM / CsmCodebooksHead
    on CPU (oracle/gen_golden.py:golden_csm_frames).  All ids and all logits bit-exact over 5 frames x 8 codebooks."""
    from oracle import csm as ocsm

    gd = _load(golden_dir, "csm_tiny_frames.npz")
    dims = ocsm.CsmDims.tiny()
    w = ocsm.synth_weights(dims, seed=int(gd["weight_seed"]))
    out = ocsm.generate_frames(w, dims, torch.from_numpy(gd["prompt_ids"]), torch.from_numpy(gd["prompt_masks"]),
                               len(gd["frames"]), page_size=int(gd["page_size"]))
    assert out["frames"] == gd["frames"].tolist()
    assert np.array_equal(torch.stack(out["cb0_logits"]).numpy(), gd["cb0_logits"])
    assert np.array_equal(torch.stack(out["depth_logits"]).numpy(), gd["depth_logits"])


def test_qwen3_tts_frame_oracle_matches_reference_golden(golden_dir):
    """BASELINE.json configs[2] / SURVEY row a24: oracle/qwen3_tts.py -- q/k RMSNorm per head before a plain RoPE,
    text-projection MLP + codec embedding + input_features as the talker input, code predictor with the head chosen by
    the largest depth position, the sum of the predictor embeddings fed back -- against the reference's own talker and
    code predictor on CPU (oracle/gen_golden.py:golden_qwen3_tts_frames).  5 frames x 6 codebooks, bit-exact."""
    from oracle import qwen3_tts as oq

    gd = _load(golden_dir, "qwen3_tts_tiny_frames.npz")
    d = oq.Qwen3TTSDims.tiny()
    w = oq.synth_weights(d, seed=int(gd["weight_seed"]))
    out = oq.generate_frames(w, d, torch.from_numpy(gd["text"]), torch.from_numpy(gd["cb0"]),
                             torch.from_numpy(gd["needs_codec"]), torch.from_numpy(gd["features"]).to(torch.bfloat16),
                             len(gd["frames"]), page_size=int(gd["page_size"]))
    assert out["frames"] == gd["frames"].tolist()
    assert np.array_equal(torch.stack(out["cb0_logits"]).numpy(), gd["cb0_logits"])
    assert np.array_equal(torch.stack(out["cp_logits"]).numpy(), gd["cp_logits"])
    full = oq.Qwen3TTSDims()         # the reference's defaults (qwen3_tts.py:113-253)
    assert (full.hidden_size, full.num_hidden_layers, full.num_code_groups, full.cp_hidden_size,
            full.cp_num_hidden_layers, full.vocab_size, full.cp_vocab_size) == (2048, 28, 16, 1024, 5, 3072, 2048)


def test_mimi_decode_oracle_matches_reference_golden(golden_dir):
    """SURVEY row a25 (CSM's vocoder): oracle/mimi.py -- split RVQ decode, learnt channel-wise x2 upsampling, the
    8-layer causal RoPE transformer with LayerScale, the SEANet decoder with causal convs / transposed convs on a
    zero left context -- against the reference's own MimiModel.decode on CPU (oracle/gen_golden.py:golden_mimi), all
    four SEANet ratios, 3 x 8 x 5 codes -> 3 x 9600 samples: latent, transformer output and waveform bit-exact."""
    from oracle import mimi as omimi

    gd = _load(golden_dir, "mimi_tiny.npz")
    cfg = omimi.MimiConfig.tiny()
    sd = omimi.synth_state_dict(cfg, int(gd["weight_seed"]))
    codes = torch.from_numpy(gd["codes"])
    with torch.no_grad():
        lat = omimi.upsample(sd, cfg, omimi.quantizer_decode(sd, cfg, codes))
        tro = omimi.transformer(sd, cfg, lat)
    assert np.array_equal(lat.numpy(), gd["latent"]) and np.array_equal(tro.numpy(), gd["transformer_out"])
    wav = omimi.decode(sd, cfg, codes)
    assert wav.shape == (3, 1, 5 * cfg.hop) and cfg.hop == 1920 and omimi.MimiConfig().hop == 1920
    assert np.array_equal(wav.numpy(), gd["wav"])


def test_qwen3_codec_streaming_decoder_oracle_matches_reference_golden(golden_dir):
    """SURVEY rows a25 / f2 (Qwen3-TTS 12 Hz codec, streaming): oracle/qwen3_codec.py against the reference's own
    Qwen3TTSTokenizerV2Decoder.forward_chunk on CPU (oracle/gen_golden.py:golden_qwen3_codec): three consecutive chunks (5, 5, 3
    frames: the 12-slot attention window wraps, the last chunk is shorter than the dilation-9 conv caches), every waveform
    and every tensor of the final cache bit-exact."""
    from oracle import qwen3_codec as oq

    gd = _load(golden_dir, "qwen3_codec_tiny.npz")
    cfg = oq.Qwen3CodecConfig.tiny()
    sd = oq.synth_state_dict(cfg, int(gd["weight_seed"]))
    cache = oq.init_cache(cfg, 2)
    for i in range(3):
        codes = torch.from_numpy(gd[f"codes{i}"])
        wav, cache = oq.forward_chunk(sd, cfg, codes, cache)
        assert wav.shape == (2, 1, codes.shape[2] * cfg.hop)
        assert np.array_equal(wav.numpy(), gd[f"wav{i}"]), i
    assert np.array_equal(cache["attention"].numpy(), gd["attention_cache"])
    assert np.array_equal(cache["position_offset"].numpy(), gd["position_offset"]) and int(cache["position_offset"][0]) == 13
    assert np.array_equal(cache["pre_conv"].numpy(), gd["pre_conv_cache"])
    for name, key in (("upsample", "upsample_conv_caches"), ("decoder_conv", "decoder_conv_caches"), ("transconv", "transconv_caches")):
        for j, t in enumerate(cache[name]):
            assert np.array_equal(t.numpy(), gd[f"{key}.{j}"]), (name, j)
    assert cfg.hop == 192 and oq.Qwen3CodecConfig().hop == 1920


def test_glm_encoder_oracle_matches_reference_golden(golden_dir):
    """oracle/glm_encoder.py against the reference's own GLMWhisperVQEncoder outputs (bf16 on CPU): ids and the last
    layer's hidden state bit for bit, for a full-length input and one with a padded tail."""
    from oracle import glm_encoder as oenc

    gd = _load(golden_dir, "glm_encoder_tiny.npz")
    d = oenc.GLMEncoderDims.tiny()
    sd = oenc.synth_state_dict(d, int(gd["weight_seed"]))
    for tag in ("full", "padded"):
        feats = torch.from_numpy(gd[f"{tag}_features"]).to(torch.bfloat16)
        mask = torch.from_numpy(gd[f"{tag}_mask"])
        ids, hidden, pooled, dist = oenc.encode(sd, d, feats, mask, return_states=True)
        assert np.array_equal(ids.numpy(), gd[f"{tag}_ids"])
        assert np.array_equal(hidden.float().numpy(), gd[f"{tag}_hidden"])
        # the per-row key bound the CUDA kernel takes IS the reference's additive mask
        am = mask[:, ::2]
        bounds = oenc.block_causal_bounds(am, d.quantize_causal_block_size)
        full = oenc.block_causal_mask(am, d.quantize_causal_block_size)[0, 0]
        T = am.shape[1]
        visible = torch.arange(T)[None, :] < bounds[:, None].long()
        assert torch.equal(full == 0, visible)
