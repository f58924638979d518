"""CPU restatement of the Qwen3-TTS talker + code-predictor frame step (BASELINE.json configs[2]: "Qwen3-TTS-1.7B
continuous batching").  TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/gen_golden.py, never
by the product path.

Follows ``vox_serve/model/qwen3_tts.py``:
  * ``Qwen3TTSAttention`` (:578-653): q / k / v / o without bias, RMSNorm over head_dim on every q and k head
    (``q_norm`` / ``k_norm``) BEFORE a plain non-interleaved RoPE at theta 1e6 (the talker's mRoPE "is essentially the
    same as standard RoPE" there); ``Qwen3TTSMLP`` (:562-575); pre-norm residual layers and a final norm for both
    stacks (:667-773);
  * talker input of a decode step (``Qwen3TTSModel.forward`` :1805-1862): ``text_projection(text_embedding(text id))``
    (a 2-layer SiLU MLP with bias, :656-664) + ``codec_embedding(cb0)`` + ``input_features`` where the latter carries
    the sum of the previous frame's predictor embeddings of codebooks 1 .. N-1 (``depth_sampling`` :1981-2004);
    ``codec_head`` without bias gives codebook 0 (:906-921);
  * code predictor (``forward_depth`` :923-944): ``small_to_mtp_projection`` (Linear with bias when the widths differ),
    its own stack, and ONE head for all rows of a call, chosen by the largest depth position: ``lm_head[max(pos) - 1]``;
  * the loop is the worker's (cuda_graph_worker.py:1058-1160): 2-row prefill [talker hidden state,
    talker.codec_embedding(cb0)] at positions 0, 1 on a zeroed per-frame cache, then 1-row decodes at position i + 1
    fed with ``code_predictor.codec_embedding[i - 1](cb_i)``.
Pinned to the reference's own modules executed on CPU: tests/golden/qwen3_tts_tiny_frames.npz
(oracle/gen_golden.py:golden_qwen3_tts_frames).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import lm_ops


@dataclass
class Qwen3TTSDims:
    # talker (Qwen3TTSTalkerConfig :205-253)
    hidden_size: int = 2048
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 6144
    vocab_size: int = 3072
    text_vocab_size: int = 151936
    text_hidden_size: int = 2048
    num_code_groups: int = 16
    # code predictor (Qwen3TTSCodePredictorConfig :113-202)
    cp_hidden_size: int = 1024
    cp_num_hidden_layers: int = 5
    cp_num_attention_heads: int = 16
    cp_num_key_value_heads: int = 8
    cp_head_dim: int = 128
    cp_intermediate_size: int = 3072
    cp_vocab_size: int = 2048
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    tts_pad_token_id: int = 151671

    @classmethod
    def tiny(cls, **kw):
        d = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                 intermediate_size=512, vocab_size=96, text_vocab_size=120, text_hidden_size=192, num_code_groups=6,
                 cp_hidden_size=128, cp_num_hidden_layers=2, cp_num_attention_heads=2, cp_num_key_value_heads=1,
                 cp_head_dim=64, cp_intermediate_size=256, cp_vocab_size=64, tts_pad_token_id=7)
        d.update(kw)
        return cls(**d)


TK, CP = "talker.model.", "talker.code_predictor.model."


def _layer(prefix: str, i: int) -> Dict[str, str]:
    p = f"{prefix}layers.{i}."
    n = {"ln1": p + "input_layernorm.weight", "ln2": p + "post_attention_layernorm.weight",
         "qn": p + "self_attn.q_norm.weight", "kn": p + "self_attn.k_norm.weight",
         "gate": p + "mlp.gate_proj.weight", "up": p + "mlp.up_proj.weight", "down": p + "mlp.down_proj.weight"}
    for x in "qkvo":
        n[x] = p + f"self_attn.{x}_proj.weight"
    return n


def _stack(w, prefix, n_layers, n_q, n_kv, D, eps, theta, h, position_ids, wrapper, kv_cache):
    t = h.shape[0]
    for i in range(n_layers):
        n = _layer(prefix, i)
        x = lm_ops.rms_norm(h, w[n["ln1"]], eps)
        q = lm_ops.rms_norm(F.linear(x, w[n["q"]]).view(-1, D), w[n["qn"]], eps).view(t, n_q, D)
        k = lm_ops.rms_norm(F.linear(x, w[n["k"]]).view(-1, D), w[n["kn"]], eps).view(t, n_kv, D)
        v = F.linear(x, w[n["v"]]).view(t, n_kv, D)
        q, k = lm_ops.apply_rope_pos_ids(q, k, position_ids, interleave=False, rope_theta=theta)
        wrapper.set_kv_cache(kv_cache[i], k, v)
        a = wrapper.run(q, kv_cache[i]).reshape(t, -1)
        h = h + F.linear(a, w[n["o"]])
        x = lm_ops.rms_norm(h, w[n["ln2"]], eps)
        h = h + F.linear(F.silu(F.linear(x, w[n["gate"]])) * F.linear(x, w[n["up"]]), w[n["down"]])
    return lm_ops.rms_norm(h, w[prefix + "norm.weight"], eps)


def talker_embeds(w, text_ids: torch.Tensor, cb0: torch.Tensor, needs_codec: torch.Tensor,
                  input_features: torch.Tensor) -> torch.Tensor:
    """rows: text id, codebook-0 id, mask[:, -1], input_features [T, H] -> [T, H]   (:1835-1853)"""
    t = F.embedding(text_ids.long(), w[TK + "text_embedding.weight"])
    t = F.linear(F.silu(F.linear(t, w["talker.text_projection.linear_fc1.weight"],
                                 w["talker.text_projection.linear_fc1.bias"])),
                 w["talker.text_projection.linear_fc2.weight"], w["talker.text_projection.linear_fc2.bias"])
    c = F.embedding(cb0.long(), w[TK + "codec_embedding.weight"])
    return torch.where(needs_codec[:, None], t + c, t) + input_features


def talker_forward(w, d: Qwen3TTSDims, inputs_embeds, position_ids, wrapper, kv_cache):
    h = _stack(w, TK, d.num_hidden_layers, d.num_attention_heads, d.num_key_value_heads, d.head_dim, d.rms_norm_eps,
               d.rope_theta, inputs_embeds, position_ids, wrapper, kv_cache)
    return F.linear(h, w["talker.codec_head.weight"]), h


def predictor_forward(w, d: Qwen3TTSDims, inputs_embeds, position_ids, wrapper, kv_cache) -> torch.Tensor:
    h = F.linear(inputs_embeds, w["talker.code_predictor.small_to_mtp_projection.weight"],
                 w["talker.code_predictor.small_to_mtp_projection.bias"])
    h = _stack(w, CP, d.cp_num_hidden_layers, d.cp_num_attention_heads, d.cp_num_key_value_heads, d.cp_head_dim,
               d.rms_norm_eps, d.rope_theta, h, position_ids, wrapper, kv_cache)
    n_heads = d.num_code_groups - 1
    idx = min(max(int(position_ids.max()), 1), n_heads) - 1                          # :936-941
    return F.linear(h, w[f"talker.code_predictor.lm_head.{idx}.weight"])


def predictor_loop_greedy(w, d: Qwen3TTSDims, hidden: torch.Tensor, cb0: int, page_size: int = 32, forced=None):
    """codebooks 1 .. N-1 of one frame for one request; returns (ids, logits [N-1, V], sum of their embeddings).
    ``forced`` (ids of codebooks 1 .. N-1): feed these back (and sum THEIR embeddings) instead of the argmax -- teacher
    forcing; the returned ids stay the oracle's own argmax of every step."""
    N = d.num_code_groups
    kv = torch.zeros(d.cp_num_hidden_layers, 1, 2, page_size, d.cp_num_key_value_heads, d.cp_head_dim,
                     dtype=hidden.dtype)
    x = torch.stack([hidden, w[TK + "codec_embedding.weight"][cb0]], dim=0)
    pre = lm_ops.PagedWrapperCPU("prefill", page_size)
    pre.plan([0, 2], [0, 1], [0], [2])
    logits = predictor_forward(w, d, x, torch.tensor([0, 1], dtype=torch.int32), pre, kv)[-1]
    ids, logs = [], []
    feat = torch.zeros(1, d.hidden_size, dtype=hidden.dtype)
    for i in range(1, N):
        logs.append(logits.float())
        tok = int(torch.argmax(logits.float()))
        ids.append(tok)
        fed = tok if forced is None else int(forced[i - 1])
        emb = w[f"{CP}codec_embedding.{i - 1}.weight"][fed][None, :]
        feat += emb                                                                  # :2002 (in-place bf16 adds)
        if i == N - 1:
            break
        dec = lm_ops.PagedWrapperCPU("decode", page_size)
        dec.plan([0, 1], [0], [i + 2])
        logits = predictor_forward(w, d, emb, torch.tensor([i + 1], dtype=torch.int32), dec, kv)[0]
    return ids, torch.stack(logs), feat


def synth_weights(d: Qwen3TTSDims, seed: int = 0, dtype=torch.bfloat16, head_scale: float = 8.0):
    """Seeded weights under the reference's state_dict names (qwen3_tts.py:707-833)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    w = {}

    def stack(prefix, n_layers, H, I, nq, nkv, D):
        for i in range(n_layers):
            n = _layer(prefix, i)
            w[n["ln1"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
            w[n["ln2"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
            w[n["qn"]] = (1.0 + rnd(D, std=0.1).float()).to(dtype)
            w[n["kn"]] = (1.0 + rnd(D, std=0.1).float()).to(dtype)
            w[n["q"]], w[n["k"]], w[n["v"]] = rnd(nq * D, H), rnd(nkv * D, H), rnd(nkv * D, H)
            w[n["o"]] = rnd(H, nq * D)
            w[n["gate"]], w[n["up"]], w[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
        w[prefix + "norm.weight"] = (1.0 + rnd(H, std=0.1).float()).to(dtype)

    H, Hc, Ht, N = d.hidden_size, d.cp_hidden_size, d.text_hidden_size, d.num_code_groups
    stack(TK, d.num_hidden_layers, H, d.intermediate_size, d.num_attention_heads, d.num_key_value_heads, d.head_dim)
    w[TK + "codec_embedding.weight"] = rnd(d.vocab_size, H, std=1.0)
    w[TK + "text_embedding.weight"] = rnd(d.text_vocab_size, Ht, std=1.0)
    w["talker.text_projection.linear_fc1.weight"], w["talker.text_projection.linear_fc1.bias"] = rnd(Ht, Ht, std=0.05), rnd(Ht, std=0.1)
    w["talker.text_projection.linear_fc2.weight"], w["talker.text_projection.linear_fc2.bias"] = rnd(H, Ht, std=0.05), rnd(H, std=0.1)
    w["talker.codec_head.weight"] = rnd(d.vocab_size, H, std=0.02 * head_scale)
    stack(CP, d.cp_num_hidden_layers, Hc, d.cp_intermediate_size, d.cp_num_attention_heads, d.cp_num_key_value_heads,
          d.cp_head_dim)
    for i in range(N - 1):
        w[f"{CP}codec_embedding.{i}.weight"] = rnd(d.cp_vocab_size, H, std=1.0)
        w[f"talker.code_predictor.lm_head.{i}.weight"] = rnd(d.cp_vocab_size, Hc, std=0.02 * head_scale)
    w["talker.code_predictor.small_to_mtp_projection.weight"] = rnd(Hc, H, std=0.05)
    w["talker.code_predictor.small_to_mtp_projection.bias"] = rnd(Hc, std=0.1)
    return w


def generate_frames(w, d: Qwen3TTSDims, prompt_text: torch.Tensor, prompt_cb0: torch.Tensor,
                    prompt_needs_codec: torch.Tensor, prompt_features: torch.Tensor, n_frames: int,
                    page_size: int = 16) -> Dict[str, List]:
    """Single request: talker prefill on the prompt rows, then ``n_frames`` frames (codebook 0 greedy from the talker,
    the rest from the predictor loop); the next talker row is text = tts_pad, cb0, mask True and
    input_features = sum of the predictor embeddings of codebooks 1 .. N-1 (:1934-1946, :2002)."""
    T0 = prompt_text.shape[0]
    n_pages = (T0 + n_frames + page_size - 1) // page_size + 1
    kv = torch.zeros(d.num_hidden_layers, n_pages, 2, page_size, d.num_key_value_heads, d.head_dim,
                     dtype=prompt_features.dtype)
    pages = list(range((T0 + page_size - 1) // page_size))
    pre = lm_ops.PagedWrapperCPU("prefill", page_size)
    pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
    logits, hidden = talker_forward(w, d, talker_embeds(w, prompt_text, prompt_cb0, prompt_needs_codec, prompt_features),
                                    torch.arange(T0, dtype=torch.int32), pre, kv)
    logits, hidden = logits[-1], hidden[-1]
    frames, cb0_logits, cp_logits, kv_len = [], [], [], T0
    for _ in range(n_frames):
        cb0_logits.append(logits.float())
        cb0 = int(torch.argmax(logits.float()))
        rest, cl, feat = predictor_loop_greedy(w, d, hidden, cb0)
        frames.append([cb0] + rest)
        cp_logits.append(cl)
        kv_len += 1
        if (kv_len + page_size - 1) // page_size > len(pages):
            pages.append(len(pages))
        dec = lm_ops.PagedWrapperCPU("decode", page_size)
        dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
        e = talker_embeds(w, torch.tensor([d.tts_pad_token_id]), torch.tensor([cb0]), torch.tensor([True]), feat)
        lg, hd = talker_forward(w, d, e, torch.tensor([kv_len - 1], dtype=torch.int32), dec, kv)
        logits, hidden = lg[0], hd[0]
    return {"frames": frames, "cb0_logits": cb0_logits, "cp_logits": cp_logits}
