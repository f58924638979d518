"""CPU oracle (test infrastructure, never on the product path) for the GLM-4-Voice speech tokenizer: the Whisper-style
VQ encoder that turns the log-mel features of a spoken prompt into ``<|audio_N|>`` token ids -- the STS prompt side of
BASELINE configs[4].

Restates ``vox_serve/encoder/glm.py``:
  * ``CausalConv1d``                              :84-107   (left padding ``dilation * (k - 1)``, then a plain Conv1d)
  * ``GLMWhisperAttention.forward``               :144-173  (q / v / out with bias, k without; SDPA with an additive mask)
  * ``GLMWhisperVQEncoderLayer.forward``          :195-214  (pre-LN residual blocks, exact-erf GELU)
  * ``GLMWhisperVQEncoder.vector_quantize``       :247-258  (``addmm(|c|^2 + |x|^2, x, c^T, alpha=-2)`` then ``min``)
  * ``get_block_causal_attention_mask``           :260-277  (causal OR same block, AND key not padding)
  * ``GLMWhisperVQEncoder.forward``               :279-323  (two convs + GELU, positions, layers, avg-pool, quantise)

Every tensor keeps the dtype the reference computes in (bf16 on the serving path, ``glm_voice.py`` loads the encoder
with the model dtype): the golden file ``tests/golden/glm_encoder_tiny.npz`` holds the reference module's own outputs
on CPU (``oracle/gen_golden.py``) and this restatement reproduces them bit for bit (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class GLMEncoderDims:
    d_model: int = 1280
    encoder_attention_heads: int = 20
    encoder_ffn_dim: int = 5120
    num_mel_bins: int = 128
    max_source_positions: int = 1500
    pooling_kernel_size: int = 4
    pooling_position: int = 16
    quantize_position: int = 16
    quantize_vocab_size: int = 16384
    quantize_causal_block_size: int = 200

    @classmethod
    def tiny(cls) -> "GLMEncoderDims":
        return cls(d_model=128, encoder_attention_heads=2, encoder_ffn_dim=256, num_mel_bins=16,
                   max_source_positions=96, pooling_kernel_size=4, pooling_position=2, quantize_position=2,
                   quantize_vocab_size=64, quantize_causal_block_size=16)


def synth_state_dict(d: GLMEncoderDims, seed: int = 0, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded weights under the reference module's parameter names (``GLMWhisperVQEncoder.state_dict()``).  The
    codebook rows are scaled like the pooled hidden states they quantise (unit-variance LayerNorm-free residual stream
    of order 1 per element), so different frames land on different codes."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    D, Fd, M = d.d_model, d.encoder_ffn_dim, d.num_mel_bins
    sd = {
        "embed_positions.weight": rnd(d.max_source_positions, D, std=0.5),
        "embed_positions2.weight": rnd(d.max_source_positions // d.pooling_kernel_size, D, std=0.5),
        "codebook.weight": rnd(d.quantize_vocab_size, D, std=1.5),
        "conv1.weight": rnd(D, M, 3, std=(3 * M) ** -0.5), "conv1.bias": rnd(D, std=0.1),
        "conv2.weight": rnd(D, D, 3, std=(3 * D) ** -0.5 * 2.0), "conv2.bias": rnd(D, std=0.1),
        "ema_count": torch.ones(d.quantize_vocab_size, dtype=dtype),
    }
    for i in range(d.quantize_position):
        p = f"layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = rnd(D, D, std=D ** -0.5)
            if n != "k_proj":
                sd[p + f"self_attn.{n}.bias"] = rnd(D, std=0.05)
        sd[p + "self_attn_layer_norm.weight"] = (1.0 + torch.randn(D, generator=g) * 0.1).to(dtype)
        sd[p + "self_attn_layer_norm.bias"] = rnd(D, std=0.05)
        sd[p + "fc1.weight"], sd[p + "fc1.bias"] = rnd(Fd, D, std=D ** -0.5), rnd(Fd, std=0.05)
        sd[p + "fc2.weight"], sd[p + "fc2.bias"] = rnd(D, Fd, std=Fd ** -0.5), rnd(D, std=0.05)
        sd[p + "final_layer_norm.weight"] = (1.0 + torch.randn(D, generator=g) * 0.1).to(dtype)
        sd[p + "final_layer_norm.bias"] = rnd(D, std=0.05)
    sd["ema_weight"] = sd["codebook.weight"].clone()
    return sd


def causal_conv1d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int) -> torch.Tensor:
    """glm.py:84-107: pad ``k - 1`` zeros on the left, then Conv1d(padding=0)."""
    k = w.shape[-1]
    x = F.pad(x.unsqueeze(2), (k - 1, 0, 0, 0)).squeeze(2)
    return F.conv1d(x, w, b, stride=stride)


def block_causal_bounds(attention_mask: torch.Tensor, block_size: int) -> torch.Tensor:
    """The reference's mask (glm.py:260-277) is ``(causal | same block) & key_not_padding``: with padding only at the
    end, row i sees exactly the keys ``j < min(end of i's block, valid length)`` -- returned per row (what the CUDA
    kernel takes as its per-row key bound)."""
    T = attention_mask.shape[-1]
    valid = int(attention_mask.reshape(-1, T)[0].sum())
    i = torch.arange(T)
    return torch.minimum((i // block_size + 1) * block_size, torch.tensor(valid)).to(torch.int32)


def block_causal_mask(attention_mask: torch.Tensor, block_size: int, dtype=torch.bfloat16) -> torch.Tensor:
    """glm.py:260-277, additive ``[B, 1, T, T]`` mask in ``dtype`` (0 where visible, finfo.min elsewhere)."""
    B, T = attention_mask.shape
    causal = torch.tril(torch.ones(1, T, T, dtype=torch.bool))
    blocks = [causal.new_ones((min(s + block_size, T) - s,) * 2) for s in range(0, T, block_size)]
    m = causal | torch.block_diag(*blocks)
    m = m & attention_mask[:, None, :].bool()
    m = m.to(dtype)
    return ((1.0 - m) * torch.finfo(dtype).min).unsqueeze(1)


def encoder_layer(h: torch.Tensor, sd: Dict[str, torch.Tensor], p: str, n_heads: int, mask: torch.Tensor) -> torch.Tensor:
    """glm.py:144-173, 195-214."""
    B, T, D = h.shape
    hd = D // n_heads
    res = h
    x = F.layer_norm(h, (D,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"])

    def heads(t):
        return t.view(B, T, n_heads, hd).transpose(1, 2).contiguous()

    q = heads(F.linear(x, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]))
    k = heads(F.linear(x, sd[p + "self_attn.k_proj.weight"]))
    v = heads(F.linear(x, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]))
    a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)
    a = a.transpose(1, 2).reshape(B, T, D)
    h = res + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    res = h
    x = F.layer_norm(h, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"])
    x = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    x = F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    return res + x


def vector_quantize(x: torch.Tensor, codebook: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """glm.py:247-258 -> (indices [rows], distances [rows, vocab])."""
    flat = x.reshape(-1, codebook.shape[1])
    c2 = torch.sum(codebook ** 2, dim=1)
    x2 = torch.sum(flat ** 2, dim=1, keepdim=True)
    dist = torch.addmm(c2 + x2, flat, codebook.t(), alpha=-2.0, beta=1.0)
    return torch.min(dist, dim=1)[1], dist


def encode(sd: Dict[str, torch.Tensor], d: GLMEncoderDims, input_features: torch.Tensor,
           attention_mask: torch.Tensor, return_states: bool = False):
    """``GLMWhisperVQEncoder.forward`` (glm.py:279-323): input_features [B, mel, frames], attention_mask [B, frames]
    -> token ids [B, frames / (2 * pooling)].  ``return_states``: also the hidden state after the last layer, the
    pooled state and the distance matrix (for tolerance-based comparisons of the CUDA path)."""
    B = input_features.shape[0]
    T = input_features.shape[-1] // 2
    am = attention_mask[:, ::2]
    mask = block_causal_mask(am, d.quantize_causal_block_size, input_features.dtype)
    x = F.gelu(causal_conv1d(input_features, sd["conv1.weight"], sd["conv1.bias"], 1))
    x = F.gelu(causal_conv1d(x, sd["conv2.weight"], sd["conv2.bias"], 2))
    h = x.permute(0, 2, 1) + sd["embed_positions.weight"][:T]
    ids = hidden_last = pooled = dist = None
    for i in range(d.quantize_position):
        h = encoder_layer(h, sd, f"layers.{i}.", d.encoder_attention_heads, mask)
        if i + 1 == d.pooling_position and d.pooling_kernel_size is not None:
            hidden_last = h
            hp = h.permute(0, 2, 1)
            if hp.shape[-1] % d.pooling_kernel_size != 0:
                hp = F.pad(hp, (0, d.pooling_kernel_size - hp.shape[-1] % d.pooling_kernel_size))
            h = F.avg_pool1d(hp, d.pooling_kernel_size).permute(0, 2, 1)
            am = am[:, :: d.pooling_kernel_size]
            mask = block_causal_mask(am, d.quantize_causal_block_size // d.pooling_kernel_size, input_features.dtype)
        if i + 1 == d.quantize_position:
            pooled = h
            idx, dist = vector_quantize(h, sd["codebook.weight"])
            ids = idx.reshape(B, h.shape[1])
    if return_states:
        return ids, hidden_last, pooled, dist
    return ids
