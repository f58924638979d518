"""The CosyVoice2 and GLM-4-Voice decoder stacks on the sm_100a engine (VERDICT r01 "missing" item 4; BASELINE.json
configs[0] and configs[4]'s LM): q / k / v bias, biased output head, embeddings-in prefill, rotation of half of every head in
(even, odd) pairs, fused gate | up weights.  Checked against the golden files produced by the reference's own
``CosyVoice2ForCausalLM`` / ``GLMVoiceForCausalLM`` on CPU (tests/golden/cosyvoice2_tiny_lm.npz, glm_voice_tiny_lm.npz):
prefill + greedy decode steps across two page boundaries, teacher-forced with the golden ids, every step's logits within
2 % of the row scale and the argmax equal unless the golden's own top-2 margin is a near-tie."""
import numpy as np
import pytest
import torch

from oracle import cosyvoice2 as ocv, glm_voice as oglm

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL = 2e-2


def _i32(x):
    return torch.tensor(x, dtype=torch.int32, device="cuda")


def _check(logits, ref, tok, st):
    ref = torch.from_numpy(ref)
    err = float((logits.float().cpu() - ref).abs().max() / ref.abs().max())
    st["max_err"] = max(st["max_err"], err)
    st["rows"] += 1
    if int(torch.argmax(logits.float())) != tok:
        top2 = torch.topk(ref, 2).values
        assert float(top2[0] - top2[1]) <= 2 * TOL * float(ref.abs().max())
        st["flips"] += 1


def _greedy(forward_prefill, forward_decode, eng, gd, T0, page):
    from vox_serve_b200 import ops

    ids, st = gd["ids"].tolist(), dict(rows=0, flips=0, max_err=0.0)
    n_pages = (T0 + page - 1) // page
    ops.plan_rows(eng.plan, _i32([0, T0]), _i32([0, n_pages]), _i32(list(range(n_pages))), _i32([T0 - (n_pages - 1) * page]),
                  1, T0, page, eng.chunk)
    logits = forward_prefill()
    torch.cuda.synchronize()
    _check(logits[0], gd["logits"][0], ids[0], st)
    kv_len = T0
    for step in range(1, len(gd["logits"])):
        kv_len += 1
        n_pages = (kv_len + page - 1) // page
        ops.plan_rows(eng.plan, None, _i32([0, n_pages]), _i32(list(range(n_pages))), _i32([kv_len - (n_pages - 1) * page]),
                      1, 1, page, eng.chunk)
        logits = forward_decode(ids[step - 1], kv_len - 1)        # teacher forcing with the golden's id
        torch.cuda.synchronize()
        if step < len(ids):
            _check(logits[0], gd["logits"][step], ids[step], st)
    return st


def test_cosyvoice2_lm_against_reference_golden(golden_dir):
    from vox_serve_b200.lm_variants import CosyVoice2LM, cosyvoice2_dims

    gd = np.load(f"{golden_dir}/cosyvoice2_tiny_lm.npz")
    od = ocv.CosyVoice2Dims.tiny()
    w = ocv.synth_weights(od, seed=int(gd["weight_seed"]))
    dims = cosyvoice2_dims(od.hidden_size, od.num_hidden_layers, od.num_attention_heads, od.num_key_value_heads,
                           od.intermediate_size, od.speech_token_size, od.rms_norm_eps, od.rope_theta)
    page, T0 = int(gd["page_size"]), int(gd["prompt_len"])
    kv = torch.zeros(dims.num_hidden_layers, 8, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF, device="cuda")
    lm = CosyVoice2LM(w, dims, kv, page, max_rows=64)
    assert not lm.engine.fused_ok                      # biased projections always take the default 8-launch layer
    emb = torch.randn(T0, od.hidden_size, generator=torch.Generator().manual_seed(int(gd["prompt_seed"]))).to(BF).cuda()
    eng = lm.engine
    st = _greedy(lambda: lm.forward_embeds(emb, torch.arange(T0, dtype=torch.int32, device="cuda"), last_rows=_i32([T0 - 1])),
                 lambda tok, pos: lm.forward_speech_ids(_i32([tok]), _i32([pos])), eng, gd, T0, page)
    print("cosyvoice2 lm vs reference golden:", st)
    assert st["max_err"] < TOL and st["flips"] <= 1, st
    full = cosyvoice2_dims()
    assert (full.hidden_size, full.num_hidden_layers, full.head_dim, full.vocab_size) == (896, 24, 64, 6564)


def test_glm_voice_lm_against_reference_golden(golden_dir):
    from vox_serve_b200.lm_variants import GLMVoiceLM, glm_voice_dims

    gd = np.load(f"{golden_dir}/glm_voice_tiny_lm.npz")
    od = oglm.GLMVoiceDims.tiny()
    w = oglm.synth_weights(od, seed=int(gd["weight_seed"]))
    dims = glm_voice_dims(od.hidden_size, od.num_layers, od.num_attention_heads, od.multi_query_group_num,
                          od.ffn_hidden_size, od.padded_vocab_size, od.layernorm_epsilon, od.rope_ratio, od.rope_theta)
    page, prompt = int(gd["page_size"]), torch.from_numpy(gd["prompt"])
    T0 = prompt.numel()
    kv = torch.zeros(dims.num_hidden_layers, 8, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF, device="cuda")
    lm = GLMVoiceLM(w, dims, kv, page, max_rows=64)
    assert dims.rotary_dim == dims.head_dim // 2 and dims.rope_interleave and not lm.engine.fused_ok
    st = _greedy(lambda: lm.forward(prompt.to(torch.int32).cuda(), torch.arange(T0, dtype=torch.int32, device="cuda"),
                                    last_rows=_i32([T0 - 1])),
                 lambda tok, pos: lm.forward(_i32([tok]), _i32([pos])), lm.engine, gd, T0, page)
    print("glm-4-voice lm vs reference golden:", st)
    assert st["max_err"] < TOL and st["flips"] <= 1, st
    full = glm_voice_dims()
    assert (full.hidden_size, full.num_hidden_layers, full.num_key_value_heads, full.vocab_size) == (4096, 40, 2, 168960)
