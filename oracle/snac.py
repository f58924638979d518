"""Oracle restatement of the SNAC decode path (``vox_serve/tokenizer/snac.py``).

* RVQ ``from_codes`` / ``decode_code``: ``snac.py:297-301, 350-357``
* ``Decoder`` / ``DecoderBlock`` / ``ResidualUnit`` / ``NoiseBlock``: ``snac.py:119-157, 160-176, 201-241``
* ``snake``: ``snac.py:252-258``;  weight-norm re-derivation ``g * v / ||v||``: ``snac.py:244-249``
  (torch ``parametrizations.weight_norm``, norm over every dim but 0)

Functional, fp32, torch CPU, reading the reference's own ``state_dict`` key names
(``...parametrizations.weight.original0/1``).  ``NoiseBlock`` draws ``torch.randn`` inside the
reference decoder; here the noise tensors are explicit inputs (one ``[B,1,T]`` per DecoderBlock).
Test infrastructure only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class SnacConfig:
    """Decode-side fields of the SNAC constructor (snac.py:362-376); defaults = snac_24khz."""
    sampling_rate: int = 24000
    encoder_dim: int = 48
    encoder_rates: Sequence[int] = (2, 4, 8, 8)
    decoder_dim: int = 1024
    decoder_rates: Sequence[int] = (8, 8, 4, 2)
    codebook_size: int = 4096
    codebook_dim: int = 8
    vq_strides: Sequence[int] = (4, 2, 1)
    latent_dim: Optional[int] = None

    @property
    def latent(self) -> int:
        return self.latent_dim or self.encoder_dim * (2 ** len(self.encoder_rates))

    @classmethod
    def tiny(cls):
        return cls(encoder_dim=4, encoder_rates=(2, 2, 2, 2), decoder_dim=64, decoder_rates=(8, 8, 4, 2),
                   codebook_size=4096, codebook_dim=8, vq_strides=(4, 2, 1))


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    dims = tuple(range(1, v.dim()))
    return v * (g / torch.linalg.vector_norm(v, ord=2, dim=dims, keepdim=True))


def _w(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    return fold_weight_norm(sd[prefix + ".parametrizations.weight.original0"].float(),
                            sd[prefix + ".parametrizations.weight.original1"].float())


def snake(x: torch.Tensor, alpha: torch.Tensor) -> torch.Tensor:
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def from_codes(sd, cfg: SnacConfig, codes: List[torch.Tensor]) -> torch.Tensor:
    z = 0.0
    for i, stride in enumerate(cfg.vq_strides):
        p = f"quantizer.quantizers.{i}"
        e = F.embedding(codes[i].long(), sd[p + ".codebook.weight"].float()).transpose(1, 2)
        zi = F.conv1d(e, _w(sd, p + ".out_proj"), sd[p + ".out_proj.bias"].float())
        z = z + zi.repeat_interleave(stride, dim=-1)
    return z


def _residual_unit(sd, p: str, x: torch.Tensor, dilation: int) -> torch.Tensor:
    c = x.shape[1]
    y = snake(x, sd[p + ".block.0.alpha"].float())
    y = F.conv1d(y, _w(sd, p + ".block.1"), sd[p + ".block.1.bias"].float(),
                 dilation=dilation, padding=3 * dilation, groups=c)
    y = snake(y, sd[p + ".block.2.alpha"].float())
    y = F.conv1d(y, _w(sd, p + ".block.3"), sd[p + ".block.3.bias"].float())
    return x + y


def noise_shapes(cfg: SnacConfig, batch: int, t_latent: int):
    """[B,1,T] shape of the randn each DecoderBlock draws."""
    out, t = [], t_latent
    for s in cfg.decoder_rates:
        t = t * s
        out.append((batch, 1, t))
    return out


def decoder(sd, cfg: SnacConfig, z: torch.Tensor, noises: Sequence[torch.Tensor]) -> torch.Tensor:
    c = z.shape[1]
    x = F.conv1d(z, _w(sd, "decoder.model.0"), sd["decoder.model.0.bias"].float(), padding=3, groups=c)
    x = F.conv1d(x, _w(sd, "decoder.model.1"), sd["decoder.model.1.bias"].float())
    li = 2
    for bi, s in enumerate(cfg.decoder_rates):
        p = f"decoder.model.{li}"
        x = snake(x, sd[p + ".block.0.alpha"].float())
        x = F.conv_transpose1d(x, _w(sd, p + ".block.1"), sd[p + ".block.1.bias"].float(),
                               stride=s, padding=math.ceil(s / 2), output_padding=s % 2)
        h = F.conv1d(x, _w(sd, p + ".block.2.linear"))
        x = x + noises[bi] * h
        for j, dil in enumerate((1, 3, 9)):
            x = _residual_unit(sd, f"{p}.block.{3 + j}", x, dil)
        li += 1
    x = snake(x, sd[f"decoder.model.{li}.alpha"].float())
    x = F.conv1d(x, _w(sd, f"decoder.model.{li + 1}"), sd[f"decoder.model.{li + 1}.bias"].float(), padding=3)
    return torch.tanh(x)


def decode(sd, cfg: SnacConfig, codes: List[torch.Tensor], noises: Sequence[torch.Tensor]) -> torch.Tensor:
    """codes: [B,T*? ] per codebook -> waveform [B,1,T_latent*prod(rates)] fp32 (snac.py:438-441)."""
    return decoder(sd, cfg, from_codes(sd, cfg, codes), noises)


def synth_state_dict(cfg: SnacConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded decoder+quantizer weights under the reference's state_dict names.  ``g`` is drawn
    independently of ``||v||`` so the weight-norm fold is exercised; scales keep activations O(1)."""
    gen = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=1.0):
        return torch.randn(*shape, generator=gen, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}

    def wn(prefix, shape, fan_in, bias_dim=None, gain=1.0):
        v = rnd(*shape)
        n = torch.linalg.vector_norm(v, dim=tuple(range(1, len(shape))), keepdim=True)
        # target per-row norm ~ gain*sqrt(rowsize/fan_in), jittered +-20 %
        rowsize = 1
        for s in shape[1:]:
            rowsize *= s
        g = gain * math.sqrt(rowsize / fan_in) * (1.0 + 0.2 * (2 * torch.rand(n.shape, generator=gen) - 1))
        sd[prefix + ".parametrizations.weight.original0"] = g
        sd[prefix + ".parametrizations.weight.original1"] = v
        if bias_dim is not None:
            sd[prefix + ".bias"] = rnd(bias_dim, std=0.05)

    L, D = cfg.latent, cfg.decoder_dim
    for i in range(len(cfg.vq_strides)):
        p = f"quantizer.quantizers.{i}"
        sd[p + ".codebook.weight"] = rnd(cfg.codebook_size, cfg.codebook_dim)
        wn(p + ".out_proj", (L, cfg.codebook_dim, 1), cfg.codebook_dim * len(cfg.vq_strides), L)
    wn("decoder.model.0", (L, 1, 7), 7, L)
    wn("decoder.model.1", (D, L, 1), L, D)
    li, cin = 2, D
    for s in cfg.decoder_rates:
        cout = cin // 2
        p = f"decoder.model.{li}"
        sd[p + ".block.0.alpha"] = 0.5 + torch.rand(1, cin, 1, generator=gen)
        wn(p + ".block.1", (cin, cout, 2 * s), 2 * cin, cout, gain=0.6)
        wn(p + ".block.2.linear", (cout, cout, 1), cout, None, gain=0.3)
        for j in range(3):
            q = f"{p}.block.{3 + j}"
            sd[q + ".block.0.alpha"] = 0.5 + torch.rand(1, cout, 1, generator=gen)
            wn(q + ".block.1", (cout, 1, 7), 7, cout)
            sd[q + ".block.2.alpha"] = 0.5 + torch.rand(1, cout, 1, generator=gen)
            wn(q + ".block.3", (cout, cout, 1), cout, cout, gain=0.5)
        li, cin = li + 1, cout
    sd[f"decoder.model.{li}.alpha"] = 0.5 + torch.rand(1, cin, 1, generator=gen)
    wn(f"decoder.model.{li + 1}", (1, cin, 7), 7 * cin, 1, gain=0.25)
    return sd
