"""Qwen3-TTS talker + code-predictor frames on the GPU (BASELINE.json configs[2]; SURVEY.md §8 row a24;
vox_serve/model/qwen3_tts.py:535-944, 1805-2004) against the golden file produced by the reference's own talker /
code-predictor modules on CPU (tests/golden/qwen3_tts_tiny_frames.npz) and against the CPU oracle (oracle/qwen3_tts.py,
pinned bit-exactly to that golden) on a ragged batch captured as one CUDA graph per frame.  What this path adds over the
Llama-shaped LM: per-head q/k RMSNorm before a plain RoPE, the text-projection + codec-embedding + input_features talker
input, a biased projection into the predictor, per-codebook predictor tables and heads, the bf16 running sum of the
predictor embeddings fed back as input_features."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import qwen3_tts as oq

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _i32(x):
    return torch.tensor(x, dtype=torch.int32, device="cuda")


def _engine(odims, weights, page, pages, max_batch, max_rows):
    from vox_serve_b200.depth_engine import Qwen3TTSDims, Qwen3TTSEngine, Qwen3TTSWeights

    dims = Qwen3TTSDims(**dataclasses.asdict(odims))
    w = Qwen3TTSWeights(weights, dims)
    kv = torch.zeros(dims.num_hidden_layers, pages, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF, device="cuda")
    return Qwen3TTSEngine(w, kv, page, max_batch=max_batch, max_rows=max_rows)


def _check(got_logits, ref_logits, got_id, ref_id, st, tol=2e-2):
    ref = ref_logits.float()
    st["max_err"] = max(st["max_err"], float((got_logits.float().cpu() - ref).abs().max() / ref.abs().max()))
    st["rows"] += 1
    if got_id != ref_id:
        top2 = torch.topk(ref, 2).values
        assert float(top2[0] - top2[1]) <= 2 * tol * float(ref.abs().max()), (got_id, ref_id)
        st["flips"] += 1


def test_qk_norm_rope_append_against_oracle():
    """The QKV tail with Qwen3's per-head q/k RMSNorm (qwen3_tts.py:603-625): reduce -> bf16 -> RMSNorm(head) -> plain
    RoPE -> q out / K,V scatter, against oracle.lm_ops on the same partials."""
    from oracle import lm_ops
    from vox_serve_b200 import ops

    T, nq, nkv, D, page = 7, 4, 2, 128, 16
    g = torch.Generator().manual_seed(0)
    parts = torch.randn(3, T, (nq + 2 * nkv) * D, generator=g)
    qn = (1 + 0.1 * torch.randn(D, generator=g)).to(BF)
    kn = (1 + 0.1 * torch.randn(D, generator=g)).to(BF)
    pos = torch.tensor([0, 1, 2, 5, 9, 30, 31], dtype=torch.int32)
    x = parts.sum(0).to(BF)
    q, k, v = x[:, :nq * D].view(T, nq, D), x[:, nq * D:(nq + nkv) * D].view(T, nkv, D), x[:, (nq + nkv) * D:].view(T, nkv, D)
    q = lm_ops.rms_norm(q.reshape(-1, D), qn, 1e-6).view(T, nq, D)
    k = lm_ops.rms_norm(k.reshape(-1, D), kn, 1e-6).view(T, nkv, D)
    q_ref, k_ref = lm_ops.apply_rope_pos_ids(q, k, pos, interleave=False, rope_theta=1e6)
    kv = torch.zeros(2, 2, page, nkv, D, dtype=BF, device="cuda")
    plan = ops.RowPlan(T, "cuda")
    ops.plan_rows(plan, _i32([0, T]), _i32([0, 1]), _i32([1]), _i32([T]), 1, T, page, ops.attn_chunk_tokens(page, nkv))
    freq = ops.rope_freq_table(D, 1.0, 1e6, False, device="cuda")
    q_out = ops.qkv_rope_append(parts.cuda(), kv, pos.cuda(), freq, plan, nq, nkv, D, q_norm=qn.cuda(), k_norm=kn.cuda(),
                                norm_eps=1e-6)
    torch.cuda.synchronize()
    dq = (q_out.float().cpu() - q_ref.float()).abs().max().item()
    dk = (kv[1, 0, :T].float().cpu() - k_ref.float()).abs().max().item()
    assert dq <= 2 * 2 ** -8 * q_ref.float().abs().max().item() and dk <= 2 * 2 ** -8 * k_ref.float().abs().max().item(), (dq, dk)
    assert torch.equal(kv[1, 1, :T].cpu(), v)                      # V is neither normalised nor rotated


def test_qwen3_tts_tiny_frames_against_reference_golden(golden_dir):
    from vox_serve_b200 import ops
    from vox_serve_b200.sampling import SamplingConfig

    gd = np.load(f"{golden_dir}/qwen3_tts_tiny_frames.npz")
    odims = oq.Qwen3TTSDims.tiny()
    weights = oq.synth_weights(odims, seed=int(gd["weight_seed"]))
    page, N = int(gd["page_size"]), odims.num_code_groups
    eng = _engine(odims, weights, page, pages=8, max_batch=2, max_rows=64)
    cfg = SamplingConfig(greedy=True)
    T0 = gd["text"].shape[0]
    n_pages = (T0 + page - 1) // page
    ops.plan_rows(eng.bb.plan, _i32([0, T0]), _i32([0, n_pages]), _i32(list(range(n_pages))),
                  _i32([T0 - (n_pages - 1) * page]), 1, T0, page, eng.bb.chunk)
    st = dict(rows=0, flips=0, max_err=0.0)
    kv_len = T0
    for f in range(len(gd["frames"])):
        keep = []
        if f == 0:
            out = eng.prefill_frame(torch.from_numpy(gd["text"]).cuda(), torch.from_numpy(gd["cb0"]).cuda(),
                                    torch.from_numpy(gd["needs_codec"]).cuda(),
                                    torch.from_numpy(gd["features"]).to(BF).cuda(),
                                    torch.arange(T0, dtype=torch.int32, device="cuda"), _i32([T0 - 1]), eng.bb.plan, cfg,
                                    keep_logits=keep)
        else:
            kv_len += 1
            n_pages = (kv_len + page - 1) // page
            ops.plan_rows(eng.bb.plan, None, _i32([0, n_pages]), _i32(list(range(n_pages))),
                          _i32([kv_len - (n_pages - 1) * page]), 1, 1, page, eng.bb.chunk)
            out = eng.decode_frame(1, _i32([kv_len - 1]), eng.bb.plan, cfg, keep_logits=keep)
        torch.cuda.synchronize()
        got, want = out[0].cpu().tolist(), gd["frames"][f].tolist()
        _check(keep[0][0], torch.from_numpy(gd["cb0_logits"][f]), got[0], want[0], st)
        for c in range(1, N):
            _check(keep[c][0], torch.from_numpy(gd["cp_logits"][f][c - 1]), got[c], want[c], st)
            if got[c] != want[c]:
                break
        if got != want:        # teacher forcing: the reference's frame and the input_features that follow from it
            eng.frame[:N, 0] = torch.tensor(want, dtype=torch.int64, device="cuda")
            eng._finish_frame(1)
    print("qwen3-tts tiny vs reference golden:", st)
    assert st["max_err"] < 2e-2 and st["flips"] <= 2, st


def test_qwen3_tts_batch_frames_one_graph_against_oracle():
    from vox_serve_b200 import ops
    from vox_serve_b200.sampling import SamplingConfig

    odims = oq.Qwen3TTSDims.tiny(hidden_size=512, num_hidden_layers=3, num_attention_heads=4, num_key_value_heads=2,
                                 head_dim=128, intermediate_size=1024, vocab_size=384, text_vocab_size=300, text_hidden_size=256,
                                 num_code_groups=8, cp_hidden_size=256, cp_num_hidden_layers=2, cp_num_attention_heads=2,
                                 cp_num_key_value_heads=1, cp_head_dim=128, cp_intermediate_size=512, cp_vocab_size=256,
                                 tts_pad_token_id=11)
    weights = oq.synth_weights(odims, seed=9)
    N, page, n_frames, H = odims.num_code_groups, 32, 4, odims.hidden_size
    g = torch.Generator().manual_seed(4)
    lens = [40, 12, 33]
    prompts = []
    for T in lens:
        text = torch.randint(0, odims.text_vocab_size, (T,), generator=g)
        cb0 = torch.randint(0, odims.vocab_size, (T,), generator=g)
        needs = torch.rand(T, generator=g) < 0.5
        feats = (torch.randn(T, H, generator=g) * 0.5).to(BF)
        prompts.append((text, cb0, needs, feats))
    ref = [oq.generate_frames(weights, odims, *p, n_frames, page_size=page) for p in prompts]
    B = len(lens)
    pages_per = [(T + n_frames + page - 1) // page + 1 for T in lens]
    eng = _engine(odims, weights, page, pages=sum(pages_per), max_batch=4, max_rows=256)
    cfg = SamplingConfig(greedy=True)
    base = np.cumsum([0] + pages_per).tolist()
    kv_len = list(lens)

    def table():
        npg = [(kv + page - 1) // page for kv in kv_len]
        return (_i32(np.cumsum([0] + npg).tolist()), _i32([base[r] + j for r in range(B) for j in range(npg[r])]),
                _i32([kv_len[r] - (npg[r] - 1) * page for r in range(B)]))

    st = dict(rows=0, flips=0, max_err=0.0)
    qo = np.cumsum([0] + lens).tolist()
    indptr, indices, last = table()
    ops.plan_rows(eng.bb.plan, _i32(qo), indptr, indices, last, B, qo[-1], page, eng.bb.chunk)
    pos = torch.cat([torch.arange(T, dtype=torch.int32) for T in lens]).cuda()
    cat = [torch.cat([p[i] for p in prompts]).cuda() for i in range(4)]
    keep = []
    out = eng.prefill_frame(cat[0], cat[1], cat[2], cat[3], pos, _i32([x - 1 for x in qo[1:]]), eng.bb.plan, cfg, keep_logits=keep)
    d_pos = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_indptr, _, d_last = [torch.zeros_like(x) for x in table()]
    d_indices = torch.zeros(sum(pages_per), dtype=torch.int32, device="cuda")
    graph = out_g = None
    for f in range(n_frames):
        torch.cuda.synchronize()
        got = out[:B].cpu().tolist()
        for r in range(B):
            want = ref[r]["frames"][f]
            if keep:
                _check(keep[0][r], ref[r]["cb0_logits"][f], got[r][0], want[0], st)
                for c in range(1, N):
                    _check(keep[c][r], ref[r]["cp_logits"][f][c - 1], got[r][c], want[c], st)
                    if got[r][c] != want[c]:
                        break
            else:
                st["flips"] += int(got[r] != want)
            eng.frame[:N, r] = torch.tensor(want, dtype=torch.int64, device="cuda")
        eng._finish_frame(B)
        if f == n_frames - 1:
            break
        kv_len = [k + 1 for k in kv_len]
        indptr, indices, last = table()
        d_pos.copy_(_i32([k - 1 for k in kv_len]))
        d_indptr.copy_(indptr), d_last.copy_(last)
        d_indices[:indices.numel()].copy_(indices)
        if f == 0:
            ops.plan_rows(eng.bb.plan, None, d_indptr, d_indices, d_last, B, B, page, eng.bb.chunk)
            keep = []
            out = eng.decode_frame(B, d_pos, eng.bb.plan, cfg, keep_logits=keep)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            n0 = ops.launch_count()
            with torch.cuda.stream(s):
                with torch.cuda.graph(graph, stream=s):
                    ops.plan_rows(eng.bb.plan, None, d_indptr, d_indices, d_last, B, B, page, eng.bb.chunk)
                    out_g = eng.decode_frame(B, d_pos, eng.bb.plan, cfg)
            torch.cuda.current_stream().wait_stream(s)
            st["graph_nodes"] = ops.launch_count() - n0
        else:
            keep = []
            graph.replay()
            out = out_g
    print("qwen3-tts batch frames vs oracle:", st)
    assert st["max_err"] < 2e-2 and st["flips"] <= 3 and st["graph_nodes"] > 80, st
