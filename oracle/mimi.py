"""CPU restatement of the Mimi DECODE path (CSM's vocoder, BASELINE.json configs[3]).
TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/gen_golden.py, never by the product path.

Follows ``vox_serve/tokenizer/mimi.py`` (``MimiModel.decode`` :2993-3018) on a plain ``{name: tensor}`` state dict with the
reference's names:
  * ``SplitResidualVectorQuantizer.decode`` (:830-836): codebook 0 through ``rvq_first``, codebooks 1.. through
    ``rvq_rest``; each = sum over its layers of ``F.embedding(codes_k, embedding_sum_k / clamp(cluster_usage_k, eps))``
    (``EuclideanCodebook.embedding`` :166-174, ``ResidualVectorQuantization.decode`` :482-490) followed by the 1x1
    ``output_proj`` without bias (:690-700); the two results are added;
  * ``_to_encoder_framerate`` -> ``ConvTrUpsample1d`` (:2272-2323): learnt channel-wise ConvTranspose1d (k = 2 s, s = 2,
    groups = C, no bias), causal: the K - S rightmost outputs are dropped (``StreamingConvTranspose1d.forward`` :2192-2215);
  * ``ProjectedTransformer`` / ``StreamingTransformerLayer`` (:1550-1896) in its stateless form: pre-LayerNorm (eps 1e-5),
    packed in_proj without bias, interleaved-pair RoPE at offset 0 with period 10 000 (``apply_rope`` :874-930), causal
    attention (the 250-step context never binds on a 20-step chunk), out_proj, LayerScale, then LayerNorm -> Linear ->
    exact GELU -> Linear -> LayerScale; no biases in the linears;
  * ``SEANetDecoder`` (:2548-2700): causal Conv1d k7 -> 4 x [ELU, causal ConvTranspose1d (k = 2 r, stride r) ,
    residual block (ELU, causal conv k3 to dim/2, ELU, conv k1 back, true skip)] -> ELU -> causal conv k3 -> 1 channel.
    Every ``StreamingConv1d`` starts from a FRESH zero state on every call (:2116-2148): each chunk is decoded with zero
    left context, which is what is reproduced here.
Pinned to the reference's own ``MimiModel`` executed on CPU: tests/golden/mimi_tiny.npz (oracle/gen_golden.py:golden_mimi).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


@dataclass
class MimiConfig:
    """``_mimi_config`` of mimi.py:21-71 (defaults = kyutai/moshiko's tokenizer)."""
    dimension: int = 512            # seanet / transformer width
    n_filters: int = 64
    ratios: Tuple[int, ...] = (8, 6, 5, 4)
    kernel_size: int = 7
    residual_kernel_size: int = 3
    last_kernel_size: int = 3
    compress: int = 2
    n_q: int = 32
    bins: int = 2048
    codebook_dim: int = 256
    num_heads: int = 8
    num_layers: int = 8
    dim_feedforward: int = 2048
    max_period: float = 10000.0
    upsample_stride: int = 2        # encoder frame rate 25 Hz / frame rate 12.5 Hz
    codebook_eps: float = 1e-5

    @classmethod
    def tiny(cls, **kw):
        d = dict(dimension=64, n_filters=8, n_q=8, bins=64, codebook_dim=32, num_heads=2, num_layers=2, dim_feedforward=128)
        d.update(kw)
        return cls(**d)

    @property
    def hop(self) -> int:
        """output samples per frame: upsample stride x prod(ratios) (1920 at the defaults)"""
        return self.upsample_stride * math.prod(self.ratios)


def _codebook(sd, prefix: str, eps: float) -> torch.Tensor:
    return sd[prefix + "embedding_sum"] / sd[prefix + "cluster_usage"].clamp(min=eps)[:, None]


def quantizer_decode(sd, cfg: MimiConfig, codes: torch.Tensor) -> torch.Tensor:
    """codes [B, K, T] -> [B, dimension, T]"""
    out = None
    for name, cols in (("rvq_first", codes[:, :1]), ("rvq_rest", codes[:, 1:])):
        q = None
        for k in range(cols.shape[1]):
            e = F.embedding(cols[:, k], _codebook(sd, f"quantizer.{name}.vq.layers.{k}._codebook.", cfg.codebook_eps))
            e = e.transpose(1, 2)                                   # "b n d -> b d n"
            q = e if q is None else q + e
        q = F.conv1d(q, sd[f"quantizer.{name}.output_proj.weight"])
        out = q if out is None else out + q
    return out


def upsample(sd, cfg: MimiConfig, x: torch.Tensor) -> torch.Tensor:
    s = cfg.upsample_stride
    y = F.conv_transpose1d(x, sd["upsample.convtr.convtr.convtr.weight"], stride=s, groups=x.shape[1])
    return y[..., : y.shape[-1] - s]                                # K - S = s trimmed on the right


def rope_interleaved(q: torch.Tensor, k: torch.Tensor, max_period: float):
    """q, k [B, H, T, D]; pairs (2j, 2j+1) rotated by t * exp(-ln(P) * 2j / D), offset 0"""
    D, T = q.shape[-1], q.shape[-2]
    freqs = torch.exp(torch.arange(D // 2, dtype=torch.float32) * (-math.log(max_period) * 2 / D))
    ang = torch.arange(T, dtype=torch.float32).view(1, 1, T, 1) * freqs
    c, s = torch.cos(ang), torch.sin(ang)

    def rot(x):
        xr, xi = x[..., 0::2].float(), x[..., 1::2].float()
        return torch.stack((xr * c - xi * s, xr * s + xi * c), dim=-1).flatten(-2).to(x.dtype)

    return rot(q), rot(k)


def transformer(sd, cfg: MimiConfig, x: torch.Tensor) -> torch.Tensor:
    """x [B, C, T] (conv layout) -> [B, C, T]"""
    x = x.transpose(1, 2)
    B, T, C = x.shape
    H = cfg.num_heads
    for i in range(cfg.num_layers):
        p = f"decoder_transformer.transformer.layers.{i}."
        h = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        qkv = F.linear(h, sd[p + "self_attn.in_projs.0.weight"]).view(B, T, 3, H, C // H).permute(2, 0, 3, 1, 4)
        q, k = rope_interleaved(qkv[0], qkv[1], cfg.max_period)
        mask = torch.ones(T, T, dtype=torch.bool).tril()
        a = F.scaled_dot_product_attention(q, k, qkv[2], mask[None, None])
        a = F.linear(a.transpose(1, 2).reshape(B, T, C), sd[p + "self_attn.out_projs.0.weight"])
        x = x + sd[p + "layer_scale_1.scale"] * a
        h = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        u = F.linear(F.gelu(F.linear(h, sd[p + "linear1.weight"])), sd[p + "linear2.weight"])
        x = x + sd[p + "layer_scale_2.scale"] * u
    return x.transpose(1, 2)


def causal_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dilation: int = 1) -> torch.Tensor:
    pad = (w.shape[-1] - 1) * dilation
    return F.conv1d(F.pad(x, (pad, 0)), w, b, dilation=dilation)


def causal_convtr(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, stride: int) -> torch.Tensor:
    y = F.conv_transpose1d(x, w, b, stride=stride)
    return y[..., : y.shape[-1] - (w.shape[-1] - stride)]


def seanet_layout(cfg: MimiConfig) -> List[Tuple[str, int]]:
    """[(kind, index in decoder.model)] in execution order (SEANetDecoder.__init__ :2625-2693, 1 residual layer)"""
    out, i = [("conv_in", 0)], 1
    for _ in cfg.ratios:
        out += [("convtr", i + 1), ("res", i + 2)]          # i = ELU
        i += 3
    out.append(("conv_out", i + 1))
    return out


def seanet_decoder(sd, cfg: MimiConfig, z: torch.Tensor) -> torch.Tensor:
    x = z
    ratios = list(cfg.ratios)
    for kind, idx in seanet_layout(cfg):
        p = f"decoder.model.{idx}."
        if kind == "conv_in":
            x = causal_conv(x, sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"])
        elif kind == "convtr":
            x = causal_convtr(F.elu(x), sd[p + "convtr.convtr.weight"], sd[p + "convtr.convtr.bias"], ratios.pop(0))
        elif kind == "res":
            h = causal_conv(F.elu(x), sd[p + "block.1.conv.conv.weight"], sd[p + "block.1.conv.conv.bias"])
            h = causal_conv(F.elu(h), sd[p + "block.3.conv.conv.weight"], sd[p + "block.3.conv.conv.bias"])
            x = x + h
        else:
            x = causal_conv(F.elu(x), sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"])
    return x


def decode(sd: Dict[str, torch.Tensor], cfg: MimiConfig, codes: torch.Tensor) -> torch.Tensor:
    """codes [B, K, T] int64 -> waveform [B, 1, T * hop] fp32 (mimi.py:2993-3018)"""
    with torch.no_grad():
        emb = upsample(sd, cfg, quantizer_decode(sd, cfg, codes.long()))
        return seanet_decoder(sd, cfg, transformer(sd, cfg, emb))


def synth_state_dict(cfg: MimiConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded decode-side weights under the reference's state_dict names (fp32).  LayerScale is set to O(1) values (the
    checkpoint's are learnt; the 0.01 initialisation would hide the transformer behind the residual path)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    D, C = cfg.codebook_dim, cfg.dimension
    for name, n in (("rvq_first", 1), ("rvq_rest", cfg.n_q - 1)):
        for k in range(n):
            p = f"quantizer.{name}.vq.layers.{k}._codebook."
            sd[p + "cluster_usage"] = torch.rand(cfg.bins, generator=g) * 3 + 0.5
            sd[p + "embedding_sum"] = rnd(cfg.bins, D) * sd[p + "cluster_usage"][:, None]
            sd[p + "_initialized"] = torch.ones(1)
        sd[f"quantizer.{name}.output_proj.weight"] = rnd(C, D, 1, std=1.0 / math.sqrt(D * cfg.n_q))
        sd[f"quantizer.{name}.input_proj.weight"] = rnd(D, C, 1, std=0.05)
    sd["upsample.convtr.convtr.convtr.weight"] = rnd(C, 1, 2 * cfg.upsample_stride, std=0.7)
    F_ = cfg.dim_feedforward
    for i in range(cfg.num_layers):
        p = f"decoder_transformer.transformer.layers.{i}."
        sd[p + "self_attn.in_projs.0.weight"] = rnd(3 * C, C, std=1.0 / math.sqrt(C))
        sd[p + "self_attn.out_projs.0.weight"] = rnd(C, C, std=1.0 / math.sqrt(C))
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = rnd(C, std=0.1, mean=1.0), rnd(C, std=0.1)
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = rnd(C, std=0.1, mean=1.0), rnd(C, std=0.1)
        sd[p + "linear1.weight"] = rnd(F_, C, std=1.0 / math.sqrt(C))
        sd[p + "linear2.weight"] = rnd(C, F_, std=1.0 / math.sqrt(F_))
        sd[p + "layer_scale_1.scale"] = rnd(C, std=0.1, mean=0.5)
        sd[p + "layer_scale_2.scale"] = rnd(C, std=0.1, mean=0.5)
    ch = cfg.n_filters * 2 ** len(cfg.ratios)
    ratios = list(cfg.ratios)
    for kind, idx in seanet_layout(cfg):
        p = f"decoder.model.{idx}."
        if kind == "conv_in":
            sd[p + "conv.conv.weight"] = rnd(ch, C, cfg.kernel_size, std=1.0 / math.sqrt(C * cfg.kernel_size))
            sd[p + "conv.conv.bias"] = rnd(ch, std=0.05)
        elif kind == "convtr":
            r = ratios.pop(0)
            sd[p + "convtr.convtr.weight"] = rnd(ch, ch // 2, 2 * r, std=1.0 / math.sqrt(2 * ch))
            sd[p + "convtr.convtr.bias"] = rnd(ch // 2, std=0.05)
            ch //= 2
        elif kind == "res":
            hid = ch // cfg.compress
            sd[p + "block.1.conv.conv.weight"] = rnd(hid, ch, cfg.residual_kernel_size, std=1.0 / math.sqrt(ch * cfg.residual_kernel_size))
            sd[p + "block.1.conv.conv.bias"] = rnd(hid, std=0.05)
            sd[p + "block.3.conv.conv.weight"] = rnd(ch, hid, 1, std=0.5 / math.sqrt(hid))
            sd[p + "block.3.conv.conv.bias"] = rnd(ch, std=0.05)
        else:
            sd[p + "conv.conv.weight"] = rnd(1, ch, cfg.last_kernel_size, std=0.5 / math.sqrt(ch * cfg.last_kernel_size))
            sd[p + "conv.conv.bias"] = rnd(1, std=0.01)
    return sd
