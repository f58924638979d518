"""CPU restatement of the CosyVoice2 speech LM (BASELINE.json configs[0]: "CosyVoice2-0.5B single prompt greedy decode
on CPU (plumbing, no GPU)").  TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/gen_golden.py, never by the
product path.

Follows ``vox_serve/model/cosyvoice2.py``:
  * ``CosyVoice2Config`` (:26-38): hidden 896, 24 layers, 14 q heads / 2 kv heads (head_dim 64), intermediate 4864,
    speech vocabulary 6561 + 3, rope theta 1e6, rms eps 1e-6;
  * ``CosyVoice2Attention`` (:122-167): q / k / v projections WITH bias, o projection without; plain rotate-half
    RoPE (``apply_rope_pos_ids(..., rope_theta=theta)``: no llama-3.1 smoothing, rope_scale 1);
  * ``CosyVoice2DecoderLayer`` / ``BackboneModel`` (:170-239): pre-norm residual blocks, final RMSNorm; the model is
    driven with ``inputs_embeds`` (text-token embeddings, the two task embeddings of ``llm_embedding``, speech-token
    embeddings of ``speech_embedding``), not ids;
  * ``CosyVoice2ForCausalLM`` (:285-315): ``llm_decoder`` = Linear(hidden, 6561 + 3) WITH bias.
The attention / norm / rope operators are the shared restatements of oracle/lm_ops.py (pinned to the installed
FlashInfer on the GPU box, tests/test_gpu_flashinfer_xcheck.py).  Pinned to the reference's own modules executed on CPU:
tests/golden/cosyvoice2_tiny_lm.npz (oracle/gen_golden.py:golden_cosyvoice2_lm).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import lm_ops


@dataclass
class CosyVoice2Dims:
    hidden_size: int = 896
    num_hidden_layers: int = 24
    num_attention_heads: int = 14
    num_key_value_heads: int = 2
    intermediate_size: int = 4864
    speech_token_size: int = 6561
    vocab_size: int = 151936            # text vocabulary (embed_tokens)
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads

    @property
    def speech_vocab(self) -> int:       # llm_decoder rows: speech tokens + 3 stop ids (cosyvoice2.py:289, 390)
        return self.speech_token_size + 3

    @classmethod
    def tiny(cls, **kw):
        d = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                 intermediate_size=512, speech_token_size=93, vocab_size=300)
        d.update(kw)
        return cls(**d)


P = "llm.model.model."


def layer_names(i: int) -> Dict[str, str]:
    p = f"{P}layers.{i}."
    n = {"ln1": p + "input_layernorm.weight", "ln2": p + "post_attention_layernorm.weight",
         "o": p + "self_attn.o_proj.weight", "gate": p + "mlp.gate_proj.weight", "up": p + "mlp.up_proj.weight",
         "down": p + "mlp.down_proj.weight"}
    for x in "qkv":
        n[x] = p + f"self_attn.{x}_proj.weight"
        n[x + "_b"] = p + f"self_attn.{x}_proj.bias"
    return n


def lm_forward(w: Dict[str, torch.Tensor], dims: CosyVoice2Dims, inputs_embeds: torch.Tensor,
               position_ids: torch.Tensor, wrapper, kv_cache: torch.Tensor) -> torch.Tensor:
    """inputs_embeds [T, H], position_ids [T] int32, kv_cache [L, pages, 2, page, Hkv, D] -> logits [T, speech_vocab]
    (cosyvoice2.py:139-167, 181-207, 220-239, 300-315)."""
    h = inputs_embeds
    t, D = h.shape[0], dims.head_dim
    for i in range(dims.num_hidden_layers):
        n = layer_names(i)
        x = lm_ops.rms_norm(h, w[n["ln1"]], dims.rms_norm_eps)
        q = F.linear(x, w[n["q"]], w[n["q_b"]]).view(t, -1, D)
        k = F.linear(x, w[n["k"]], w[n["k_b"]]).view(t, -1, D)
        v = F.linear(x, w[n["v"]], w[n["v_b"]]).view(t, -1, D)
        q, k = lm_ops.apply_rope_pos_ids(q, k, position_ids, rope_theta=dims.rope_theta)
        wrapper.set_kv_cache(kv_cache[i], k, v)
        a = wrapper.run(q, kv_cache[i]).reshape(t, -1)
        h = h + F.linear(a, w[n["o"]])
        x = lm_ops.rms_norm(h, w[n["ln2"]], dims.rms_norm_eps)
        h = h + F.linear(F.silu(F.linear(x, w[n["gate"]])) * F.linear(x, w[n["up"]]), w[n["down"]])
    h = lm_ops.rms_norm(h, w[P + "norm.weight"], dims.rms_norm_eps)
    return F.linear(h, w["llm_decoder.weight"], w["llm_decoder.bias"])


def synth_weights(dims: CosyVoice2Dims, seed: int = 0, dtype=torch.bfloat16,
                  head_scale: float = 8.0) -> Dict[str, torch.Tensor]:
    """Seeded weights under the reference's state_dict names (cosyvoice2.py:242-292)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    H, I, D = dims.hidden_size, dims.intermediate_size, dims.head_dim
    hq, hkv = dims.num_attention_heads * D, dims.num_key_value_heads * D
    w = {P + "embed_tokens.weight": rnd(dims.vocab_size, H, std=1.0)}
    for i in range(dims.num_hidden_layers):
        n = layer_names(i)
        w[n["ln1"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
        w[n["ln2"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
        w[n["q"]], w[n["k"]], w[n["v"]] = rnd(hq, H), rnd(hkv, H), rnd(hkv, H)
        w[n["q_b"]], w[n["k_b"]], w[n["v_b"]] = rnd(hq, std=0.3), rnd(hkv, std=0.3), rnd(hkv, std=0.3)
        w[n["o"]] = rnd(H, hq)
        w[n["gate"]], w[n["up"]], w[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
    w[P + "norm.weight"] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
    w["llm.model.lm_head.weight"] = rnd(dims.vocab_size, H)          # unused by the path (cosyvoice2.py:247)
    w["llm_embedding.weight"] = rnd(2, H, std=1.0)
    w["llm_decoder.weight"] = rnd(dims.speech_vocab, H, std=0.02 * head_scale)
    w["llm_decoder.bias"] = rnd(dims.speech_vocab, std=0.1)
    w["speech_embedding.weight"] = rnd(dims.speech_vocab, H, std=1.0)
    return w


def greedy_decode(w, dims: CosyVoice2Dims, prompt_embeds: torch.Tensor, n_steps: int, page_size: int = 16,
                  stop_ids=None) -> Dict[str, List]:
    """Single-prompt greedy decode, the plumbing of configs[0]: prefill with ``inputs_embeds`` (masks all True),
    then feed ``speech_embedding(id)`` of the sampled token at the next position; pages are appended as the sequence
    grows (worker/base.py:210-360 bookkeeping for one request)."""
    T0 = prompt_embeds.shape[0]
    n_pages = (T0 + n_steps + page_size - 1) // page_size + 1
    kv = torch.zeros(dims.num_hidden_layers, n_pages, 2, page_size, dims.num_key_value_heads, dims.head_dim,
                     dtype=prompt_embeds.dtype)
    pages = list(range((T0 + page_size - 1) // page_size))
    pre = lm_ops.PagedWrapperCPU("prefill", page_size)
    pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
    logits = lm_forward(w, dims, prompt_embeds, torch.arange(T0, dtype=torch.int32), pre, kv)[-1:]
    ids, logs, kv_len = [], [logits[0].float()], T0
    stop = set(stop_ids) if stop_ids is not None else set(range(dims.speech_token_size, dims.speech_vocab))
    for _ in range(n_steps):
        tok = int(torch.argmax(logits[0].float()))
        ids.append(tok)
        if tok in stop:
            break
        kv_len += 1
        if (kv_len + page_size - 1) // page_size > len(pages):
            pages.append(len(pages))
        dec = lm_ops.PagedWrapperCPU("decode", page_size)
        dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
        emb = F.embedding(torch.tensor([tok]), w["speech_embedding.weight"])
        logits = lm_forward(w, dims, emb, torch.tensor([kv_len - 1], dtype=torch.int32), dec, kv)
        logs.append(logits[0].float())
    return {"ids": ids, "logits": logs}
