"""Oracle restatement of the Orpheus adapter's arithmetic (``vox_serve/model/orpheus.py``).

* decoder layer / backbone / lm_head: ``orpheus.py:125-221`` (attention ``:81-111``, MLP ``:46-48``)
* forward wrapper (codebook dim in/out): ``orpheus.py:398-417``
* sampling + repetition cache + stop rule: ``orpheus.py:419-477``
* audio-id mapping and 7-token frame de-interleave: ``orpheus.py:479-507``

Functional style over a flat ``{hf_name: tensor}`` weight dict (HF Llama parameter names), torch CPU.
Every ``nn.Linear`` output, RMSNorm output, RoPE output, attention output and residual add is
rounded to the weight dtype exactly where the reference's module graph rounds.
Test infrastructure only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import lm_ops, sampler as osampler


@dataclass
class OrpheusDims:
    hidden_size: int = 3072
    num_hidden_layers: int = 28
    num_attention_heads: int = 24
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 8192
    vocab_size: int = 156940
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_factor: float = 32.0
    low_freq_factor: float = 1.0
    high_freq_factor: float = 4.0
    old_context_len: int = 8192
    stop_token_id: int = 128258          # orpheus.py:258
    audio_id_base: int = 128266          # 128256 + 10, orpheus.py:479-481
    max_tokens: int = 1200               # orpheus.py:310-316

    @classmethod
    def tiny(cls, **kw):
        d = dict(hidden_size=768, num_hidden_layers=2, num_attention_heads=6, num_key_value_heads=2,
                 head_dim=128, intermediate_size=1024, vocab_size=10 + 7 * 64,
                 stop_token_id=3, audio_id_base=10)
        d.update(kw)
        return cls(**d)


def layer_names(i: int) -> Dict[str, str]:
    p = f"model.layers.{i}."
    return {
        "ln1": p + "input_layernorm.weight", "ln2": p + "post_attention_layernorm.weight",
        "q": p + "self_attn.q_proj.weight", "k": p + "self_attn.k_proj.weight",
        "v": p + "self_attn.v_proj.weight", "o": p + "self_attn.o_proj.weight",
        "gate": p + "mlp.gate_proj.weight", "up": p + "mlp.up_proj.weight",
        "down": p + "mlp.down_proj.weight",
    }


def lm_forward(w: Dict[str, torch.Tensor], dims: OrpheusDims, input_ids: torch.Tensor,
               position_ids: torch.Tensor, wrapper, kv_cache: torch.Tensor,
               return_hidden: bool = False) -> torch.Tensor:
    """input_ids [T] (codebook dim already removed), position_ids [T] int32,
    kv_cache [L, pages, 2, page, Hkv, D].  Returns logits [T, V] in the weight dtype."""
    h = F.embedding(input_ids.long(), w["model.embed_tokens.weight"])
    t = h.shape[0]
    for i in range(dims.num_hidden_layers):
        n = layer_names(i)
        x = lm_ops.rms_norm(h, w[n["ln1"]], dims.rms_norm_eps)
        q = F.linear(x, w[n["q"]]).view(t, -1, dims.head_dim)
        k = F.linear(x, w[n["k"]]).view(t, -1, dims.head_dim)
        v = F.linear(x, w[n["v"]]).view(t, -1, dims.head_dim)
        q, k = lm_ops.apply_rope_pos_ids(
            q, k, position_ids, rope_scale=dims.rope_factor, rope_theta=dims.rope_theta,
            low_freq_factor=dims.low_freq_factor, high_freq_factor=dims.high_freq_factor,
            old_context_len=dims.old_context_len)
        wrapper.set_kv_cache(kv_cache[i], k, v)
        a = wrapper.run(q, kv_cache[i]).reshape(t, -1)
        h = h + F.linear(a, w[n["o"]])
        x = lm_ops.rms_norm(h, w[n["ln2"]], dims.rms_norm_eps)
        g = F.silu(F.linear(x, w[n["gate"]])) * F.linear(x, w[n["up"]])
        h = h + F.linear(g, w[n["down"]])
    h = lm_ops.rms_norm(h, w["model.norm.weight"], dims.rms_norm_eps)
    if return_hidden:
        return h
    return F.linear(h, w["lm_head.weight"])


def sampling_step(logits: torch.Tensor, cfg, rep_cache: Optional[torch.Tensor],
                  generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """logits [B, 1, V] -> ids [B, 1] int64; mutates rep_cache [B, W, 1, V] (orpheus.py:419-447).

    Greedy is exact.  Stochastic branches draw with torch.multinomial from the oracle's
    ``filtered_probs`` (distribution-level restatement; see oracle/sampler.py)."""
    if rep_cache is not None:
        logits = osampler.apply_repetition_penalty(logits, rep_cache, cfg.repetition_penalty)
    flat = logits.reshape(-1, logits.shape[-1])
    if osampler.strategy(cfg) == "greedy":
        ids = osampler.greedy(flat)
    else:
        ids = torch.multinomial(osampler.filtered_probs(flat, cfg), 1, generator=generator)[:, 0]
    ids = ids.view(logits.shape[0], logits.shape[1])
    if rep_cache is not None:
        osampler.update_repetition_cache(rep_cache, ids, cfg.repetition_window)
    return ids


def audio_codes_from_window(token_ids: torch.Tensor, dims: OrpheusDims) -> List[torch.Tensor]:
    """[B, 28, 1] LM ids -> SNAC code lists [B,4], [B,8], [B,16] (orpheus.py:483-500)."""
    mf = (token_ids.reshape(-1, 4, 7).long() - dims.audio_id_base) % 4096
    return [mf[:, :, 0], mf[:, :, [1, 4]].reshape(-1, 8), mf[:, :, [2, 3, 5, 6]].reshape(-1, 16)]


def synth_weights(dims: OrpheusDims, seed: int = 0, dtype=torch.bfloat16,
                  lm_head_scale: float = 8.0, planted: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """Seeded N(0, 0.02) weights at the given dims; RMSNorm weights near 1;
    lm_head scaled up so greedy top-1/top-2 margins sit well above bf16 noise.

    ``planted`` (embedding std, e.g. 2.0): "confident model" weights for free-running greedy parity.  With i.i.d.
    weights the top-1/top-2 gap of 157 k logits is exponentially distributed, so some row of a 32 x 64 run always
    lands within rounding noise of a tie whatever the seed.  Here ``lm_head[perm[v]] = planted_gain * embed[v]``
    for a seeded permutation: the row of the *next* token is aligned with the current token's embedding, which the
    residual stream still carries at the output (the 28 layers add ~14x its energy on top, so every logit depends
    on the whole network; the logits themselves are compared against the oracle at the usual tolerance).  The top-1
    logit then stands clear of the runner-up by a large fraction of its own value, with or without the repetition
    penalty."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=0.02):
        # (row chunks bound the fp32 temporary of the 157 k x 3072 matrices)
        out = torch.empty(*shape, dtype=dtype)
        flat = out.view(shape[0], -1) if len(shape) > 1 else out.view(-1, 1)
        step = max(1, (1 << 25) // max(1, flat.shape[1]))
        for r in range(0, flat.shape[0], step):
            blk = flat[r:r + step]
            blk.copy_(torch.randn(blk.shape, generator=g, dtype=torch.float32) * std)
        return out

    H, I = dims.hidden_size, dims.intermediate_size
    hq, hkv = dims.num_attention_heads * dims.head_dim, dims.num_key_value_heads * dims.head_dim
    w = {"model.embed_tokens.weight": rnd(dims.vocab_size, H, std=1.0 if planted is None else float(planted))}
    for i in range(dims.num_hidden_layers):
        n = layer_names(i)
        w[n["ln1"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
        w[n["ln2"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
        w[n["q"]], w[n["k"]], w[n["v"]] = rnd(hq, H), rnd(hkv, H), rnd(hkv, H)
        w[n["o"]] = rnd(H, hq)
        w[n["gate"]], w[n["up"]], w[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
    w["model.norm.weight"] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
    if planted is None:
        w["lm_head.weight"] = rnd(dims.vocab_size, H, std=0.02 * lm_head_scale)
    else:
        perm = torch.randperm(dims.vocab_size, generator=g)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(dims.vocab_size)
        gain = 32.0 / (float(planted) ** 2 * H) * 8.0          # planted logit of order 32 (see docstring)
        w["lm_head.weight"] = (w["model.embed_tokens.weight"][inv].float() * gain).to(dtype)
    return w
