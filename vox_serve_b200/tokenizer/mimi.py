"""Mimi decode path on the sm_100a kernels (drop-in for the decode side of ``vox_serve/tokenizer/mimi.py``:
``MimiDecoder(...).decode(codes [B, K, T]) -> [B, 1, T * 1920]``, mimi.py:3024-3090 / ``MimiModel.decode`` :2993-3018).

CSM's vocoder (``vox_serve/model/csm.py:771-785``): the reference decodes every 10-frame chunk independently with zero
left context (``StreamingConv1d.forward`` builds a fresh zero state per call, mimi.py:2116-2148), so
the decoder here is stateless as well.  Only the decode side exists (split RVQ codebooks + output projections, the learnt
x2 upsampling, the decoder transformer, the SEANet decoder); the encoder is prompt-side and not on the serving path.

At load time the codebooks are materialised once (``embedding_sum / clamp(cluster_usage, eps)``, which the reference
caches on first use, mimi.py:166-174), the transposed-conv weights are re-packed per output phase, and everything is kept
in fp32 like the reference module.  Activations stay in the ``[B, C, T]`` conv layout end to end: the transformer's Linear
layers are 1x1 convolutions there, LayerNorm / attention index it directly (csrc/mimi.cu).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from .._lib import VoxB200Error, call
from . import _tc

F32 = torch.float32


@dataclass
class MimiConfig:
    """``_mimi_config`` (mimi.py:21-71)."""
    dimension: int = 512
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
    upsample_stride: int = 2
    codebook_eps: float = 1e-5
    sample_rate: int = 24000

    @property
    def hop(self) -> int:
        return self.upsample_stride * math.prod(self.ratios)


def seanet_layout(cfg: MimiConfig) -> List[Tuple[str, int]]:
    """[(kind, index in decoder.model)] in execution order (SEANetDecoder.__init__, mimi.py:2625-2693)."""
    out, i = [("conv_in", 0)], 1
    for _ in cfg.ratios:
        out += [("convtr", i + 1), ("res", i + 2)]
        i += 3
    out.append(("conv_out", i + 1))
    return out


def synthetic_state_dict(cfg: MimiConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded decode-side weights under the checkpoint's own key names (for ``csm-synthetic`` models: no checkpoint can be
    fetched offline).  Scales keep every stage's activations O(1)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    C, D, Fd, sd = cfg.dimension, cfg.codebook_dim, cfg.dim_feedforward, {}
    for name, n in (("rvq_first", 1), ("rvq_rest", cfg.n_q - 1)):
        for k in range(n):
            p = f"quantizer.{name}.vq.layers.{k}._codebook."
            sd[p + "cluster_usage"] = torch.rand(cfg.bins, generator=g) + 0.5
            sd[p + "embedding_sum"] = rnd(cfg.bins, D) * sd[p + "cluster_usage"][:, None]
        sd[f"quantizer.{name}.output_proj.weight"] = rnd(C, D, 1, std=1.0 / math.sqrt(D * cfg.n_q))
    sd["upsample.convtr.convtr.convtr.weight"] = rnd(C, 1, 2 * cfg.upsample_stride, std=0.7)
    for i in range(cfg.num_layers):
        p = f"decoder_transformer.transformer.layers.{i}."
        sd[p + "self_attn.in_projs.0.weight"] = rnd(3 * C, C, std=C ** -0.5)
        sd[p + "self_attn.out_projs.0.weight"] = rnd(C, C, std=C ** -0.5)
        sd[p + "linear1.weight"], sd[p + "linear2.weight"] = rnd(Fd, C, std=C ** -0.5), rnd(C, Fd, std=Fd ** -0.5)
        for n_ in ("norm1", "norm2"):
            sd[p + n_ + ".weight"], sd[p + n_ + ".bias"] = rnd(C, std=0.1, mean=1.0), rnd(C, std=0.1)
        sd[p + "layer_scale_1.scale"], sd[p + "layer_scale_2.scale"] = rnd(C, std=0.1, mean=0.5), rnd(C, std=0.1, mean=0.5)
    ch, ratios = cfg.n_filters * 2 ** len(cfg.ratios), list(cfg.ratios)
    for kind, idx in seanet_layout(cfg):
        p = f"decoder.model.{idx}."
        if kind == "conv_in":
            sd[p + "conv.conv.weight"] = rnd(ch, C, cfg.kernel_size, std=(C * cfg.kernel_size) ** -0.5)
            sd[p + "conv.conv.bias"] = rnd(ch, std=0.05)
        elif kind == "convtr":
            r = ratios.pop(0)
            sd[p + "convtr.convtr.weight"], sd[p + "convtr.convtr.bias"] = rnd(ch, ch // 2, 2 * r, std=(2 * ch) ** -0.5), rnd(ch // 2, std=0.05)
            ch //= 2
        elif kind == "res":
            hid, k = ch // cfg.compress, cfg.residual_kernel_size
            sd[p + "block.1.conv.conv.weight"], sd[p + "block.1.conv.conv.bias"] = rnd(hid, ch, k, std=(ch * k) ** -0.5), rnd(hid, std=0.05)
            sd[p + "block.3.conv.conv.weight"], sd[p + "block.3.conv.conv.bias"] = rnd(ch, hid, 1, std=0.5 * hid ** -0.5), rnd(ch, std=0.05)
        else:
            k = cfg.last_kernel_size
            sd[p + "conv.conv.weight"], sd[p + "conv.conv.bias"] = rnd(1, ch, k, std=0.5 * (ch * k) ** -0.5), rnd(1, std=0.01)
    return sd


class MimiDecoder:
    def __init__(self, model_repo: str = "kyutai/moshiko-pytorch-bf16",
                 model_path: str = "tokenizer-e351c8d8-checkpoint125.safetensors", num_codebooks: int = 32,
                 mimi_config: Optional[MimiConfig] = None, device="cuda", state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.cfg = mimi_config or MimiConfig()
        self.device = torch.device(device)
        self.num_codebooks = num_codebooks
        self.w: Dict[str, torch.Tensor] = {}
        self.loaded = False
        if state_dict is None:
            path = os.path.join(model_repo, model_path) if os.path.isdir(model_repo) else model_repo
            if not os.path.isfile(path):
                raise VoxB200Error("no network in this build: pass state_dict=... or a local safetensors path as model_repo")
            from safetensors.torch import load_file

            state_dict = load_file(path)
        self.load_state_dict(state_dict)

    @property
    def sample_rate(self) -> int:
        return self.cfg.sample_rate

    def eval(self):
        return self

    # ---- loading ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        cfg, dev, w = self.cfg, self.device, {}

        def f(name):
            return sd[name].to(F32)

        emb = []
        for name, n in (("rvq_first", 1), ("rvq_rest", cfg.n_q - 1)):
            for k in range(n):
                p = f"quantizer.{name}.vq.layers.{k}._codebook."
                emb.append(f(p + "embedding_sum") / f(p + "cluster_usage").clamp(min=cfg.codebook_eps)[:, None])
            w[f"{name}.proj"] = f(f"quantizer.{name}.output_proj.weight")[:, :, 0]
        w["codebooks"] = torch.stack(emb)                                   # [n_q, bins, D]
        w["upsample"] = f("upsample.convtr.convtr.convtr.weight")[:, 0, :]   # [C, 2 s]
        for i in range(cfg.num_layers):
            p = f"decoder_transformer.transformer.layers.{i}."
            for short, name in (("in", "self_attn.in_projs.0.weight"), ("out", "self_attn.out_projs.0.weight"),
                                ("n1w", "norm1.weight"), ("n1b", "norm1.bias"), ("n2w", "norm2.weight"), ("n2b", "norm2.bias"),
                                ("l1", "linear1.weight"), ("l2", "linear2.weight"), ("s1", "layer_scale_1.scale"),
                                ("s2", "layer_scale_2.scale")):
                w[f"t{i}.{short}"] = f(p + name)
        self.layout = seanet_layout(cfg)
        for kind, idx in self.layout:
            p = f"decoder.model.{idx}."
            if kind in ("conv_in", "conv_out"):
                wt = f(p + "conv.conv.weight")
                w[f"d{idx}.w"], w[f"d{idx}.b"] = wt.reshape(wt.shape[0], -1), f(p + "conv.conv.bias")
            elif kind == "convtr":
                wt = f(p + "convtr.convtr.weight")                        # [Cin, Cout, 2 s]
                cin, cout, k = wt.shape
                s = k // 2
                # packed[r][co][tap * Cin + ci] = W[ci][co][r + tap * s]
                w[f"d{idx}.w"] = wt.view(cin, cout, 2, s).permute(3, 1, 2, 0).reshape(s, cout, 2 * cin)
                w[f"d{idx}.b"] = f(p + "convtr.convtr.bias")
            else:
                w1, w3 = f(p + "block.1.conv.conv.weight"), f(p + "block.3.conv.conv.weight")
                w[f"d{idx}.w1"], w[f"d{idx}.b1"] = w1.reshape(w1.shape[0], -1), f(p + "block.1.conv.conv.bias")
                w[f"d{idx}.w3"], w[f"d{idx}.b3"] = w3.reshape(w3.shape[0], -1), f(p + "block.3.conv.conv.bias")
        self.w = {k: v.contiguous().to(dev) for k, v in w.items()}
        # tensor-core copies (tf32 hi/lo tiles, csrc/snac_mma.cu) of the wide convolutions: key -> packed tiles
        self.tc: Dict[str, torch.Tensor] = {}
        kinds = {idx: kd for kd, idx in self.layout}
        for k, v in self.w.items():
            t = None
            if k in ("rvq_first.proj", "rvq_rest.proj") or (k[0] == "t" and k.split(".")[-1] in ("in", "out", "l1", "l2")):
                t = _tc.pack_conv(v, v.shape[1], 1)
            elif k[0] == "d" and k.endswith((".w", ".w1", ".w3")):
                kind = kinds[int(k[1:].split(".")[0])]
                if kind == "convtr":
                    t = _tc.pack_convtr(v, v.shape[2] // 2)
                else:
                    ks = {"conv_in": cfg.kernel_size, "conv_out": cfg.last_kernel_size}.get(
                        kind, cfg.residual_kernel_size if k.endswith(".w1") else 1)
                    t = _tc.pack_conv(v, v.shape[1] // ks, ks)
            if t is not None:
                self.tc[k] = t
        self.loaded = True
        return self

    # ---- kernels ------------------------------------------------------------------------------------------
    def _conv(self, y, x, key, bias, B, cin, cout, T, ksize=1, dil=1, epi=0, resid=None, scale=None, elu_in=False):
        """``key``: name of the weight in self.w; wide layers run on the tcgen05 kernel (zero left context: every chunk is
        decoded on its own), the rest on the fp32 SIMT kernel."""
        p = lambda t: None if t is None else t.data_ptr()            # noqa: E731
        tc = self.tc.get(key)
        if tc is not None:
            if elu_in and ksize > 1:       # ELU once, not once per tap in the tensor-core kernel's loaders
                xa = torch.empty_like(x)
                call("vb_codec_activate", xa.data_ptr(), x.data_ptr(), None, None, 1, B, cin, T, ops._stream())
                x, elu_in = xa, False
            call("vb_codec_conv_tc", y.data_ptr(), x.data_ptr(), tc.data_ptr(), p(bias), p(resid), p(scale), None, None, None, epi,
                 int(elu_in), B, cin, cout, T, ksize, dil, ops._stream())
        else:
            call("vb_mimi_conv", y.data_ptr(), x.data_ptr(), self.w[key].data_ptr(), p(bias), p(resid), p(scale), epi, int(elu_in), B,
                 cin, cout, T, ksize, dil, ops._stream())
        return y

    @torch.no_grad()
    def decode(self, codes: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """codes [B, K, T] (any int dtype) -> waveform [B, 1, T * hop] fp32 (mimi.py:2993-3018).  ``taps`` (a dict)
        receives the intermediate ``latent`` (after upsampling) and ``transformer_out`` tensors for stage-wise tests."""
        if not self.loaded:
            raise VoxB200Error("Mimi weights not loaded")
        cfg, w, dev, st = self.cfg, self.w, self.device, ops._stream()
        if not codes.is_cuda:
            raise VoxB200Error("MimiDecoder.decode needs CUDA tensors (there is no CPU path)")
        c = codes.to(torch.int64).contiguous()
        B, K, T = c.shape
        if K > cfg.n_q:
            raise VoxB200Error(f"{K} codebooks given, the quantizer has {cfg.n_q}")
        C, D = cfg.dimension, cfg.codebook_dim
        e = dict(dtype=F32, device=dev)
        # ---- split RVQ decode: rvq_first(codebook 0) + rvq_rest(codebooks 1..) ----
        z = torch.empty(B, D, T, **e)
        call("vb_mimi_codes_sum", z.data_ptr(), c.data_ptr(), w["codebooks"].data_ptr(), B, K, 0, 1, cfg.bins, D, T, st)
        x = self._conv(torch.empty(B, C, T, **e), z, "rvq_first.proj", None, B, D, C, T)
        if K > 1:
            z2 = torch.empty(B, D, T, **e)
            call("vb_mimi_codes_sum", z2.data_ptr(), c.data_ptr(), w["codebooks"].data_ptr(), B, K, 1, K, cfg.bins, D, T, st)
            x = self._conv(torch.empty(B, C, T, **e), z2, "rvq_rest.proj", None, B, D, C, T, epi=1, resid=x)
        # ---- learnt channel-wise x2 upsampling (12.5 Hz -> 25 Hz) ----
        s = cfg.upsample_stride
        u = torch.empty(B, C, T * s, **e)
        call("vb_mimi_upsample", u.data_ptr(), x.data_ptr(), w["upsample"].data_ptr(), B, C, T, s, st)
        x, T = u, T * s
        if taps is not None:
            taps["latent"] = x
        if T > 64:
            raise VoxB200Error(f"chunks of {T // s} frames exceed the attention kernel's 64 positions: decode in shorter chunks")
        # ---- decoder transformer (pre-LN, RoPE, causal, LayerScale; all Linear layers are 1x1 convs here) ----
        H, Fd = cfg.num_heads, cfg.dim_feedforward
        for i in range(cfg.num_layers):
            h = torch.empty(B, C, T, **e)
            call("vb_mimi_layernorm", h.data_ptr(), x.data_ptr(), w[f"t{i}.n1w"].data_ptr(), w[f"t{i}.n1b"].data_ptr(), B, C, T,
                 1e-5, st)
            qkv = self._conv(torch.empty(B, 3 * C, T, **e), h, f"t{i}.in", None, B, C, 3 * C, T)
            a = torch.empty(B, C, T, **e)
            call("vb_mimi_attention", a.data_ptr(), qkv.data_ptr(), B, C, H, T, float(cfg.max_period), st)
            x = self._conv(torch.empty(B, C, T, **e), a, f"t{i}.out", None, B, C, C, T, epi=2, resid=x, scale=w[f"t{i}.s1"])
            call("vb_mimi_layernorm", h.data_ptr(), x.data_ptr(), w[f"t{i}.n2w"].data_ptr(), w[f"t{i}.n2b"].data_ptr(), B, C, T,
                 1e-5, st)
            m = self._conv(torch.empty(B, Fd, T, **e), h, f"t{i}.l1", None, B, C, Fd, T, epi=3)
            x = self._conv(torch.empty(B, C, T, **e), m, f"t{i}.l2", None, B, Fd, C, T, epi=2, resid=x, scale=w[f"t{i}.s2"])
        if taps is not None:
            taps["transformer_out"] = x
        # ---- SEANet decoder ----
        ch = cfg.n_filters * 2 ** len(cfg.ratios)
        ratios = list(cfg.ratios)
        for kind, idx in self.layout:
            if kind == "conv_in":
                x = self._conv(torch.empty(B, ch, T, **e), x, f"d{idx}.w", w[f"d{idx}.b"], B, C, ch, T, ksize=cfg.kernel_size)
            elif kind == "convtr":
                r = ratios.pop(0)
                y = torch.empty(B, ch // 2, T * r, **e)
                tcw = self.tc.get(f"d{idx}.w")
                if tcw is not None:
                    call("vb_codec_convtr_tc", y.data_ptr(), x.data_ptr(), tcw.data_ptr(), w[f"d{idx}.b"].data_ptr(), None, None, None,
                         1, B, ch, ch // 2, T, r, st)
                else:
                    call("vb_mimi_convtr", y.data_ptr(), x.data_ptr(), w[f"d{idx}.w"].data_ptr(), w[f"d{idx}.b"].data_ptr(), 1, B, ch,
                         ch // 2, T, r, st)
                x, ch, T = y, ch // 2, T * r
            elif kind == "res":
                hid = ch // cfg.compress
                h = self._conv(torch.empty(B, hid, T, **e), x, f"d{idx}.w1", w[f"d{idx}.b1"], B, ch, hid, T,
                               ksize=cfg.residual_kernel_size, elu_in=True)
                x = self._conv(torch.empty(B, ch, T, **e), h, f"d{idx}.w3", w[f"d{idx}.b3"], B, hid, ch, T, epi=1, resid=x,
                               elu_in=True)
            else:
                x = self._conv(torch.empty(B, 1, T, **e), x, f"d{idx}.w", w[f"d{idx}.b"], B, ch, 1, T,
                               ksize=cfg.last_kernel_size, elu_in=True)
        return x
