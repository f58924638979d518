"""CPU restatement of the GLM-4-Voice decoder LM (BASELINE.json configs[4]: "GLM-4-Voice-9B STS").  TEST
INFRASTRUCTURE ONLY: imported by tests/ and oracle/gen_golden.py, never by the product path.

Follows ``vox_serve/model/glm_voice.py``:
  * ``GLMVoiceConfig`` (:22-54): hidden 4096, 40 layers, 32 q heads / 2 kv groups (head_dim 128), ffn 13696, padded
    vocabulary 168 960, rms eps 3.90625e-8, rope_ratio 1;
  * ``GLMVoiceAttention`` (:104-163): ONE fused ``query_key_value`` Linear with bias, split [Hq*D | Hkv*D | Hkv*D];
    RoPE on the first ``head_dim // 2`` dims, INTERLEAVED pairs, theta 10 000, rope_scale = rope_ratio; ``dense``
    output projection without bias;
  * ``GLMVoiceMLP`` (:85-101): fused ``dense_h_to_4h`` (2 * ffn rows: gate half then up half), ``silu(x0) * x1``,
    ``dense_4h_to_h``;
  * pre-norm residual layers, ``final_layernorm``, ``output_layer`` without bias (:166-278); driven with
    ``inputs_embeds`` (``embedding.word_embeddings``).
Pinned to the reference's own modules executed on CPU: tests/golden/glm_voice_tiny_lm.npz
(oracle/gen_golden.py:golden_glm_voice_lm).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import lm_ops


@dataclass
class GLMVoiceDims:
    hidden_size: int = 4096
    num_layers: int = 40
    num_attention_heads: int = 32
    multi_query_group_num: int = 2
    ffn_hidden_size: int = 13696
    padded_vocab_size: int = 168960
    layernorm_epsilon: float = 3.90625e-08
    rope_ratio: float = 1.0
    rope_theta: float = 10000.0

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @classmethod
    def tiny(cls, **kw):
        d = dict(hidden_size=512, num_layers=2, num_attention_heads=4, multi_query_group_num=2, ffn_hidden_size=640,
                 padded_vocab_size=320)
        d.update(kw)
        return cls(**d)


P = "transformer.encoder.layers."


def layer_names(i: int) -> Dict[str, str]:
    p = f"{P}{i}."
    return {"ln1": p + "input_layernorm.weight", "ln2": p + "post_attention_layernorm.weight",
            "qkv": p + "self_attention.query_key_value.weight", "qkv_b": p + "self_attention.query_key_value.bias",
            "o": p + "self_attention.dense.weight", "gu": p + "mlp.dense_h_to_4h.weight",
            "down": p + "mlp.dense_4h_to_h.weight"}


def lm_forward(w: Dict[str, torch.Tensor], dims: GLMVoiceDims, inputs_embeds: torch.Tensor,
               position_ids: torch.Tensor, wrapper, kv_cache: torch.Tensor) -> torch.Tensor:
    """inputs_embeds [T, H] -> logits [T, padded_vocab] (glm_voice.py:123-163, 177-203, 216-233, 260-278)."""
    h = inputs_embeds
    t, D = h.shape[0], dims.head_dim
    nq, nkv = dims.num_attention_heads * D, dims.multi_query_group_num * D
    for i in range(dims.num_layers):
        n = layer_names(i)
        x = lm_ops.rms_norm(h, w[n["ln1"]], dims.layernorm_epsilon)
        q, k, v = F.linear(x, w[n["qkv"]], w[n["qkv_b"]]).split([nq, nkv, nkv], dim=-1)
        q, k, v = q.reshape(t, -1, D), k.reshape(t, -1, D), v.reshape(t, -1, D)
        q, k = lm_ops.apply_rope_pos_ids(q, k, position_ids, rotary_dim=D // 2, interleave=True,
                                         rope_scale=dims.rope_ratio, rope_theta=dims.rope_theta)
        wrapper.set_kv_cache(kv_cache[i], k, v)
        a = wrapper.run(q, kv_cache[i]).reshape(t, -1)
        h = h + F.linear(a, w[n["o"]])
        x = lm_ops.rms_norm(h, w[n["ln2"]], dims.layernorm_epsilon)
        g, u = torch.chunk(F.linear(x, w[n["gu"]]), 2, dim=-1)
        h = h + F.linear(F.silu(g) * u, w[n["down"]])
    h = lm_ops.rms_norm(h, w["transformer.encoder.final_layernorm.weight"], dims.layernorm_epsilon)
    return F.linear(h, w["transformer.output_layer.weight"])


def synth_weights(dims: GLMVoiceDims, seed: int = 0, dtype=torch.bfloat16, head_scale: float = 8.0):
    """Seeded weights under the reference's state_dict names (glm_voice.py:104-288)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    H, I, D = dims.hidden_size, dims.ffn_hidden_size, dims.head_dim
    qkv = H + 2 * D * dims.multi_query_group_num
    w = {"transformer.embedding.word_embeddings.weight": rnd(dims.padded_vocab_size, H, std=1.0)}
    for i in range(dims.num_layers):
        n = layer_names(i)
        w[n["ln1"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
        w[n["ln2"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
        w[n["qkv"]], w[n["qkv_b"]] = rnd(qkv, H), rnd(qkv, std=0.3)
        w[n["o"]] = rnd(H, H)
        w[n["gu"]], w[n["down"]] = rnd(2 * I, H), rnd(H, I)
    w["transformer.encoder.final_layernorm.weight"] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
    w["transformer.output_layer.weight"] = rnd(dims.padded_vocab_size, H, std=0.02 * head_scale)
    return w


def greedy_decode(w, dims: GLMVoiceDims, prompt_ids: torch.Tensor, n_steps: int, page_size: int = 16) -> Dict[str, List]:
    """Single-request greedy decode over the paged CPU wrapper: prefill with the prompt's word embeddings, then feed
    the embedding of each sampled id (glm_voice.py:517-560 forward; worker/base.py:210-360 page bookkeeping)."""
    emb_w = w["transformer.embedding.word_embeddings.weight"]
    T0 = prompt_ids.shape[0]
    n_pages = (T0 + n_steps + page_size - 1) // page_size + 1
    kv = torch.zeros(dims.num_layers, n_pages, 2, page_size, dims.multi_query_group_num, dims.head_dim,
                     dtype=emb_w.dtype)
    pages = list(range((T0 + page_size - 1) // page_size))
    pre = lm_ops.PagedWrapperCPU("prefill", page_size)
    pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
    logits = lm_forward(w, dims, F.embedding(prompt_ids.long(), emb_w), torch.arange(T0, dtype=torch.int32), pre,
                        kv)[-1:]
    ids, logs, kv_len = [], [logits[0].float()], T0
    for _ in range(n_steps):
        tok = int(torch.argmax(logits[0].float()))
        ids.append(tok)
        kv_len += 1
        if (kv_len + page_size - 1) // page_size > len(pages):
            pages.append(len(pages))
        dec = lm_ops.PagedWrapperCPU("decode", page_size)
        dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
        logits = lm_forward(w, dims, F.embedding(torch.tensor([tok]), emb_w),
                            torch.tensor([kv_len - 1], dtype=torch.int32), dec, kv)
        logs.append(logits[0].float())
    return {"ids": ids, "logits": logs}
