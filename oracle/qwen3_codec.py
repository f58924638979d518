"""CPU restatement of the Qwen3-TTS 12 Hz codec DECODER in its streaming form (SURVEY rows a25 / f2;
``vox_serve/tokenizer/qwen3_codec.py``: ``Qwen3TTSTokenizerV2Decoder.forward_chunk`` :1541-1667 with ``init_cache`` :1381-1539).
TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/gen_golden.py, never by the product path.

One call decodes a chunk of frames ``codes [B, 16, T]`` into ``T * 1920`` samples and advances a per-request cache:
  * split RVQ decode (:1144-1304): codebook 0 through ``rvq_first``, codebooks 1.. through ``rvq_rest``; every codebook row is
    ``embedding_sum / clamp(cluster_usage, eps)``; each group has its own bias-free 1x1 ``output_proj``;
  * ``pre_conv``: causal Conv1d k3 with a 2-sample left-context cache (:239-340);
  * ``pre_transformer`` (:516-977): input_proj (Linear + bias) -> 8 pre-norm layers [RMSNorm, attention with rotate-half RoPE at
    the running position offset over a FIXED 72-slot K/V cache that is shifted left by the chunk length and attended in full
    -- slots that were never written hold zeros and still take softmax mass, exactly as the reference's zero-initialised
    cache does (:573-655) --, LayerScale, RMSNorm, SiLU-gated MLP, LayerScale] -> RMSNorm -> output_proj (Linear + bias);
  * two upsampling stages (:343-470): ConvTranspose1d with kernel == stride (no overlap, no cache) + ConvNeXt block
    (depthwise causal k7 with a 6-sample cache, LayerNorm eps 1e-6, Linear-GELU-Linear, gamma, residual);
  * decoder (:980-1142): causal conv k7 (cache 6) -> 4 x [SnakeBeta -> causal ConvTranspose1d k = 2s with a ONE-sample input
    cache (the previous chunk's last activated input is prepended, the first ``s`` and the last ``s`` raw outputs are dropped)
    -> 3 residual units (SnakeBeta, causal k7 conv with dilation 1 / 3 / 9 and its cache, SnakeBeta, 1x1 conv, + input)]
    -> SnakeBeta -> causal conv k7 (cache 6) -> clamp(-1, 1).
Everything is fp32 (``init_cache(..., torch.float32)``, ``Qwen3TTSDecoder(dtype=torch.float32)`` :1797-1800).

Pinned to the reference's own module executed on CPU: tests/golden/qwen3_codec_tiny.npz
(oracle/gen_golden.py:golden_qwen3_codec): three consecutive chunks, waveform and the final cache.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


@dataclass
class Qwen3CodecConfig:
    """``Qwen3TTSTokenizerV2DecoderConfig`` (:88-113) fields the decode path reads."""
    latent_dim: int = 1024
    codebook_dim: int = 512
    codebook_size: int = 2048
    decoder_dim: int = 1536
    hidden_size: int = 512
    intermediate_size: int = 1024
    layer_scale_initial_scale: float = 0.01
    head_dim: int = 64
    num_attention_heads: int = 16
    num_hidden_layers: int = 8
    num_key_value_heads: int = 16
    num_quantizers: int = 16
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    sliding_window: int = 72
    upsample_rates: Tuple[int, ...] = (8, 5, 4, 3)
    upsampling_ratios: Tuple[int, ...] = (2, 2)
    codebook_eps: float = 1e-5

    @classmethod
    def tiny(cls, **kw):
        d = dict(latent_dim=64, codebook_dim=32, codebook_size=64, decoder_dim=128, hidden_size=32, intermediate_size=64,
                 head_dim=8, num_attention_heads=4, num_hidden_layers=2, num_key_value_heads=2, num_quantizers=4,
                 sliding_window=12, upsample_rates=(4, 3, 2, 2), upsampling_ratios=(2, 2))
        d.update(kw)
        return cls(**d)

    @property
    def hop(self) -> int:
        """samples per frame: prod(upsample_rates) * prod(upsampling_ratios) (1920 at the defaults)"""
        return math.prod(self.upsample_rates) * math.prod(self.upsampling_ratios)


# ---------------------------------------------------------------------------------------------------------------------
# cache
# ---------------------------------------------------------------------------------------------------------------------
def init_cache(cfg: Qwen3CodecConfig, B: int) -> Dict[str, object]:
    """The state tensors of ``Qwen3TTSDecoderCache`` (:34-85) -- the work / output buffers of the reference are scratch."""
    z = lambda *s: torch.zeros(*s, dtype=torch.float32)          # noqa: E731
    c = {"attention": z(B, cfg.num_hidden_layers, cfg.num_key_value_heads, cfg.sliding_window, 2 * cfg.head_dim),
         "position_offset": torch.zeros(B, dtype=torch.long),
         "pre_conv": z(B, cfg.codebook_dim, 2),
         "upsample": [z(B, cfg.latent_dim, 6) for _ in cfg.upsampling_ratios],
         "decoder_conv": [z(B, cfg.latent_dim, 6)], "transconv": []}
    ch = cfg.decoder_dim
    for _ in cfg.upsample_rates:
        c["transconv"].append(z(B, ch, 1))
        ch //= 2
        c["decoder_conv"] += [z(B, ch, 6 * d) for d in (1, 3, 9)]
    c["decoder_conv"].append(z(B, ch, 6))
    return c


# ---------------------------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------------------------
def _codebook(sd, p: str, eps: float) -> torch.Tensor:
    return sd[p + "embedding_sum"] / sd[p + "cluster_usage"].clamp(min=eps)[:, None]


def quantizer_decode(sd, cfg: Qwen3CodecConfig, codes: torch.Tensor) -> torch.Tensor:
    """codes [B, K, T] -> [B, codebook_dim, T] (:1204-1211, 1256-1261, 1298-1304)"""
    out = None
    for name, cols in (("rvq_first", codes[:, :1]), ("rvq_rest", codes[:, 1:])):
        q = torch.zeros([1])[0]
        for k in range(cols.shape[1]):
            e = F.embedding(cols[:, k], _codebook(sd, f"quantizer.{name}.vq.layers.{k}._codebook.", cfg.codebook_eps))
            q = q + e.transpose(1, 2)
        q = F.conv1d(q, sd[f"quantizer.{name}.output_proj.weight"])
        out = q if out is None else out + q
    return out


def causal_conv_chunk(x, w, b, cache, dilation: int = 1, groups: int = 1):
    """``CausalConvNet.forward_chunk`` (:274-340): [cache | x] -> conv; the cache keeps the last (k - 1) * dilation inputs."""
    pad = (w.shape[-1] - 1) * dilation
    if pad == 0:
        return F.conv1d(x, w, b, groups=groups), cache
    full = torch.cat([cache, x], dim=2)
    L = x.shape[2]
    new_cache = x[:, :, -pad:].clone() if L >= pad else torch.cat([cache[:, :, L:], x], dim=2)
    return F.conv1d(full, w, b, dilation=dilation, groups=groups), new_cache


def transconv_chunk(x, w, b, cache, stride: int):
    """``CausalTransConvNet.forward_chunk`` (:359-397), kernel = 2 * stride."""
    L = x.shape[2]
    raw = F.conv_transpose1d(torch.cat([cache, x], dim=2), w, b, stride=stride)
    return raw[:, :, stride:stride + L * stride].contiguous(), x[:, :, -1:].clone()


def snake_beta(x, alpha, beta):
    """:1004-1018"""
    a, bt = torch.exp(alpha)[None, :, None], torch.exp(beta)[None, :, None]
    return x + (1.0 / (bt + 0.000000001)) * torch.pow(torch.sin(x * a), 2)


def rms_norm(x, w, eps):
    """:713-718"""
    v = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x.float() * torch.rsqrt(v + eps)).to(x.dtype)


def _rotate_half(x):
    return torch.cat((-x[..., x.shape[-1] // 2:], x[..., : x.shape[-1] // 2]), dim=-1)


def attention_chunk(sd, cfg: Qwen3CodecConfig, p: str, x, cos, sin, kv_cache):
    """``DecoderAttention.forward_chunk`` (:573-655): returns (output, new cache [B, Hkv, W, 2 D])."""
    B, T, _ = x.shape
    H, Hkv, D, W = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.sliding_window
    q = F.linear(x, sd[p + "q_proj.weight"]).view(B, T, H, D).transpose(1, 2)
    k = F.linear(x, sd[p + "k_proj.weight"]).view(B, T, Hkv, D).transpose(1, 2)
    v = F.linear(x, sd[p + "v_proj.weight"]).view(B, T, Hkv, D).transpose(1, 2)
    c, s = cos.unsqueeze(1), sin.unsqueeze(1)
    q, k = q * c + _rotate_half(q) * s, k * c + _rotate_half(k) * s
    new = kv_cache.clone()
    if T < W:
        new[:, :, :-T] = kv_cache[:, :, T:]
        new[:, :, -T:, :D], new[:, :, -T:, D:] = k, v
    else:
        new[:, :, :, :D], new[:, :, :, D:] = k[:, :, -W:], v[:, :, -W:]
    fk, fv = new[..., :D], new[..., D:]
    if H // Hkv > 1:
        fk, fv = fk.repeat_interleave(H // Hkv, dim=1), fv.repeat_interleave(H // Hkv, dim=1)
    kv_len = fk.shape[2]
    allowed = torch.arange(kv_len)[None, :] <= (kv_len - T + torch.arange(T))[:, None]
    mask = torch.zeros(T, kv_len, dtype=x.dtype).masked_fill_(~allowed, float("-inf"))
    o = F.scaled_dot_product_attention(q, fk, fv, attn_mask=mask)
    return F.linear(o.transpose(1, 2).contiguous().view(B, T, -1), sd[p + "o_proj.weight"]), new


def transformer_chunk(sd, cfg: Qwen3CodecConfig, x, cache) -> torch.Tensor:
    """x [B, T, latent] -> [B, T, latent]; updates cache["attention"] / ["position_offset"] (:914-977)"""
    P = "pre_transformer."
    B, T, _ = x.shape
    h = F.linear(x, sd[P + "input_proj.weight"], sd[P + "input_proj.bias"])
    pos = torch.arange(T)[None, :] + cache["position_offset"].view(-1, 1)
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, cfg.head_dim, 2, dtype=torch.float32) / cfg.head_dim))
    fr = (inv[None, :, None].expand(pos.shape[0], -1, 1) @ pos[:, None, :].float()).transpose(1, 2)
    emb = torch.cat((fr, fr), dim=-1)
    cos, sin = emb.cos(), emb.sin()
    att = cache["attention"].clone()
    for i in range(cfg.num_hidden_layers):
        L = f"{P}layers.{i}."
        a, att[:, i] = attention_chunk(sd, cfg, L + "self_attn.", rms_norm(h, sd[L + "input_layernorm.weight"], cfg.rms_norm_eps),
                                       cos, sin, att[:, i])
        h = h + sd[L + "self_attn_layer_scale.scale"] * a
        m = rms_norm(h, sd[L + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        m = F.linear(F.silu(F.linear(m, sd[L + "mlp.gate_proj.weight"])) * F.linear(m, sd[L + "mlp.up_proj.weight"]),
                     sd[L + "mlp.down_proj.weight"])
        h = h + sd[L + "mlp_layer_scale.scale"] * m
    cache["attention"] = att
    cache["position_offset"] = cache["position_offset"] + T
    h = rms_norm(h, sd[P + "norm.weight"], cfg.rms_norm_eps)
    return F.linear(h, sd[P + "output_proj.weight"], sd[P + "output_proj.bias"])


def convnext_chunk(sd, p: str, x, cache):
    """:434-468"""
    C = x.shape[1]
    h, cache = causal_conv_chunk(x, sd[p + "dwconv.conv.weight"], sd[p + "dwconv.conv.bias"], cache, groups=C)
    h = F.layer_norm(h.permute(0, 2, 1), (C,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    h = F.linear(F.gelu(F.linear(h, sd[p + "pwconv1.weight"], sd[p + "pwconv1.bias"])), sd[p + "pwconv2.weight"],
                 sd[p + "pwconv2.bias"])
    return x + (sd[p + "gamma"] * h).permute(0, 2, 1), cache


def forward_chunk(sd, cfg: Qwen3CodecConfig, codes: torch.Tensor, cache: Dict[str, object]):
    """codes [B, num_quantizers, T] -> (wav [B, 1, T * hop], cache) -- the cache dict is updated and returned (:1541-1667)"""
    assert codes.shape[1] == cfg.num_quantizers
    with torch.no_grad():
        h = quantizer_decode(sd, cfg, codes)
        h, cache["pre_conv"] = causal_conv_chunk(h, sd["pre_conv.conv.weight"], sd["pre_conv.conv.bias"], cache["pre_conv"])
        h = transformer_chunk(sd, cfg, h.transpose(1, 2), cache).permute(0, 2, 1)
        for j, f in enumerate(cfg.upsampling_ratios):
            h = F.conv_transpose1d(h, sd[f"upsample.{j}.0.conv.weight"], sd[f"upsample.{j}.0.conv.bias"], stride=f).contiguous()
            h, cache["upsample"][j] = convnext_chunk(sd, f"upsample.{j}.1.", h, cache["upsample"][j])
        dc, ci = cache["decoder_conv"], 0
        h, dc[ci] = causal_conv_chunk(h, sd["decoder.0.conv.weight"], sd["decoder.0.conv.bias"], dc[ci])
        ci += 1
        for bi, rate in enumerate(cfg.upsample_rates):
            p = f"decoder.{bi + 1}.block."
            h = snake_beta(h, sd[p + "0.alpha"], sd[p + "0.beta"])
            h, cache["transconv"][bi] = transconv_chunk(h, sd[p + "1.conv.weight"], sd[p + "1.conv.bias"], cache["transconv"][bi],
                                                        rate)
            for u, dil in enumerate((1, 3, 9)):
                q = f"{p}{u + 2}."
                r = snake_beta(h, sd[q + "act1.alpha"], sd[q + "act1.beta"])
                r, dc[ci] = causal_conv_chunk(r, sd[q + "conv1.conv.weight"], sd[q + "conv1.conv.bias"], dc[ci], dilation=dil)
                ci += 1
                r = snake_beta(r, sd[q + "act2.alpha"], sd[q + "act2.beta"])
                h = F.conv1d(r, sd[q + "conv2.conv.weight"], sd[q + "conv2.conv.bias"]) + h
        n = len(cfg.upsample_rates) + 1
        h = snake_beta(h, sd[f"decoder.{n}.alpha"], sd[f"decoder.{n}.beta"])
        h, dc[ci] = causal_conv_chunk(h, sd[f"decoder.{n + 1}.conv.weight"], sd[f"decoder.{n + 1}.conv.bias"], dc[ci])
        return h.clamp(min=-1, max=1), cache


# ---------------------------------------------------------------------------------------------------------------------
# seeded weights under the reference module's parameter names
# ---------------------------------------------------------------------------------------------------------------------
def synth_state_dict(cfg: Qwen3CodecConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    dq, C, Lt, H = cfg.codebook_dim // 2, cfg.codebook_dim, cfg.latent_dim, cfg.hidden_size
    for name, n in (("rvq_first", 1), ("rvq_rest", cfg.num_quantizers - 1)):
        for k in range(n):
            p = f"quantizer.{name}.vq.layers.{k}._codebook."
            sd[p + "cluster_usage"] = torch.rand(cfg.codebook_size, generator=g) * 3 + 0.5
            sd[p + "embedding_sum"] = rnd(cfg.codebook_size, dq) * sd[p + "cluster_usage"][:, None]
        sd[f"quantizer.{name}.input_proj.weight"] = rnd(dq, C, 1, std=0.05)
        sd[f"quantizer.{name}.output_proj.weight"] = rnd(C, dq, 1, std=1.0 / math.sqrt(dq * cfg.num_quantizers))
    sd["pre_conv.conv.weight"], sd["pre_conv.conv.bias"] = rnd(Lt, C, 3, std=1.0 / math.sqrt(3 * C)), rnd(Lt, std=0.05)
    P = "pre_transformer."
    sd[P + "input_proj.weight"], sd[P + "input_proj.bias"] = rnd(H, Lt, std=Lt ** -0.5), rnd(H, std=0.05)
    sd[P + "output_proj.weight"], sd[P + "output_proj.bias"] = rnd(Lt, H, std=H ** -0.5), rnd(Lt, std=0.05)
    sd[P + "norm.weight"] = rnd(H, std=0.1, mean=1.0)
    nh, nkv, D, I = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size
    for i in range(cfg.num_hidden_layers):
        L = f"{P}layers.{i}."
        sd[L + "self_attn.q_proj.weight"], sd[L + "self_attn.k_proj.weight"] = rnd(nh * D, H, std=H ** -0.5), rnd(nkv * D, H, std=H ** -0.5)
        sd[L + "self_attn.v_proj.weight"], sd[L + "self_attn.o_proj.weight"] = rnd(nkv * D, H, std=H ** -0.5), rnd(H, nh * D, std=(nh * D) ** -0.5)
        sd[L + "mlp.gate_proj.weight"], sd[L + "mlp.up_proj.weight"] = rnd(I, H, std=H ** -0.5), rnd(I, H, std=H ** -0.5)
        sd[L + "mlp.down_proj.weight"] = rnd(H, I, std=I ** -0.5)
        sd[L + "input_layernorm.weight"], sd[L + "post_attention_layernorm.weight"] = rnd(H, std=0.1, mean=1.0), rnd(H, std=0.1, mean=1.0)
        sd[L + "self_attn_layer_scale.scale"], sd[L + "mlp_layer_scale.scale"] = rnd(H, std=0.1, mean=0.5), rnd(H, std=0.1, mean=0.5)
    for j, f in enumerate(cfg.upsampling_ratios):
        sd[f"upsample.{j}.0.conv.weight"], sd[f"upsample.{j}.0.conv.bias"] = rnd(Lt, Lt, f, std=Lt ** -0.5), rnd(Lt, std=0.05)
        q = f"upsample.{j}.1."
        sd[q + "dwconv.conv.weight"], sd[q + "dwconv.conv.bias"] = rnd(Lt, 1, 7, std=7 ** -0.5), rnd(Lt, std=0.05)
        sd[q + "norm.weight"], sd[q + "norm.bias"] = rnd(Lt, std=0.1, mean=1.0), rnd(Lt, std=0.1)
        sd[q + "pwconv1.weight"], sd[q + "pwconv1.bias"] = rnd(4 * Lt, Lt, std=Lt ** -0.5), rnd(4 * Lt, std=0.05)
        sd[q + "pwconv2.weight"], sd[q + "pwconv2.bias"] = rnd(Lt, 4 * Lt, std=(4 * Lt) ** -0.5), rnd(Lt, std=0.05)
        sd[q + "gamma"] = rnd(Lt, std=0.1, mean=0.5)
    ch = cfg.decoder_dim
    sd["decoder.0.conv.weight"], sd["decoder.0.conv.bias"] = rnd(ch, Lt, 7, std=(7 * Lt) ** -0.5), rnd(ch, std=0.05)
    for bi, rate in enumerate(cfg.upsample_rates):
        p = f"decoder.{bi + 1}.block."
        sd[p + "0.alpha"], sd[p + "0.beta"] = rnd(ch, std=0.3), rnd(ch, std=0.3)
        sd[p + "1.conv.weight"], sd[p + "1.conv.bias"] = rnd(ch, ch // 2, 2 * rate, std=(2 * ch) ** -0.5), rnd(ch // 2, std=0.05)
        ch //= 2
        for u in range(3):
            q = f"{p}{u + 2}."
            sd[q + "act1.alpha"], sd[q + "act1.beta"] = rnd(ch, std=0.3), rnd(ch, std=0.3)
            sd[q + "act2.alpha"], sd[q + "act2.beta"] = rnd(ch, std=0.3), rnd(ch, std=0.3)
            sd[q + "conv1.conv.weight"], sd[q + "conv1.conv.bias"] = rnd(ch, ch, 7, std=(7 * ch) ** -0.5), rnd(ch, std=0.05)
            sd[q + "conv2.conv.weight"], sd[q + "conv2.conv.bias"] = rnd(ch, ch, 1, std=0.5 * ch ** -0.5), rnd(ch, std=0.05)
    n = len(cfg.upsample_rates) + 1
    sd[f"decoder.{n}.alpha"], sd[f"decoder.{n}.beta"] = rnd(ch, std=0.3), rnd(ch, std=0.3)
    sd[f"decoder.{n + 1}.conv.weight"], sd[f"decoder.{n + 1}.conv.bias"] = rnd(1, ch, 7, std=0.05 * (7 * ch) ** -0.5), rnd(1, std=0.01)
    return sd
