"""Qwen3-TTS 12 Hz codec decoder in its streaming form on the sm_100a kernels (drop-in for the decode side of
``vox_serve/tokenizer/qwen3_codec.py``: ``Qwen3TTSDecoder.init_cache`` :1865-1885, ``decode_chunk`` :1887-1903, i.e.
``Qwen3TTSTokenizerV2Decoder.forward_chunk`` :1541-1667 over a ``Qwen3TTSDecoderCache`` :34-85).

``decode_chunk(codes [B, 16, T], cache) -> (wav [B, 1, T * 1920], cache)``: every chunk continues from the per-request state
in ``cache`` -- the 72-slot K/V window of the 8 transformer layers with its position offset, the left-context caches of every
causal convolution and the one-sample input caches of the four transposed convolutions.  The cache is updated in place and
holds only state tensors: the reference's work / output buffers exist to make ITS torch ops CUDA-graph safe and have no
counterpart here (every kernel reads the cache as left context directly).  ``Qwen3TTSDecoderCache`` derives from
``tokenizer.base.DecoderCache``, so the worker can stack / split per-request caches exactly like the reference's.

Weights stay fp32 (``Qwen3TTSDecoder(dtype=torch.float32)``, :1797-1800); at load the codebooks are materialised, q / k / v
weights concatenated, transposed-conv weights re-packed per output phase and the SnakeBeta parameters turned into the two
per-channel tables the kernels use (csrc/codec.cu).  Only the decode side exists (the Mimi-based encoder is prompt-side).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from .._lib import VoxB200Error, call
from . import _tc
from .base import DecoderCache

F32 = torch.float32


@dataclass
class Qwen3CodecConfig:
    """``Qwen3TTSTokenizerV2DecoderConfig`` (qwen3_codec.py:88-113)."""
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
    sample_rate: int = 24000

    @property
    def hop(self) -> int:
        return math.prod(self.upsample_rates) * math.prod(self.upsampling_ratios)


@dataclass
class Qwen3TTSDecoderCache(DecoderCache):
    """The state fields of the reference's cache (qwen3_codec.py:34-85); batch is the leading dimension of every tensor."""
    attention_cache: Optional[torch.Tensor] = None          # [B, layers, Hkv, window, 2 * head_dim]
    position_offset: Optional[torch.Tensor] = None          # [B] int64
    pre_conv_cache: Optional[torch.Tensor] = None           # [B, codebook_dim, 2]
    upsample_conv_caches: List[torch.Tensor] = field(default_factory=list)      # 2 x [B, latent, 6]
    decoder_conv_caches: List[torch.Tensor] = field(default_factory=list)       # 14 x [B, C, 6 * dilation]
    transconv_caches: List[torch.Tensor] = field(default_factory=list)          # 4 x [B, C, 1]


class Qwen3TTSDecoder:
    def __init__(self, model_repo: str = "Qwen/Qwen3-TTS-Tokenizer-12Hz", device="cuda", dtype: torch.dtype = F32,
                 config: Optional[Qwen3CodecConfig] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        if dtype != F32:
            raise VoxB200Error("the codec decoder computes in fp32, like the reference's default")
        self.cfg = config or Qwen3CodecConfig()
        self.device = torch.device(device)
        if state_dict is None:
            raise VoxB200Error(f"'{model_repo}': no network in this build; pass state_dict= (decoder.* keys of the checkpoint)")
        self.load_state_dict(state_dict)

    @property
    def sample_rate(self) -> int:
        return self.cfg.sample_rate

    # ---- loading ------------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = False):
        cfg, dev, w = self.cfg, self.device, {}
        if any(k.startswith("decoder.quantizer.") for k in sd):           # the full tokenizer checkpoint nests the decoder
            sd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}

        def f(name):
            return sd[name].to(F32)

        def snake(prefix):
            return torch.exp(f(prefix + "alpha")), 1.0 / (torch.exp(f(prefix + "beta")) + 0.000000001)

        def packed_convtr(wt, two_taps: bool):
            cin, cout, k = wt.shape
            if two_taps:           # kernel 2 s: packed[r][co][tap * Cin + ci] = W[ci][co][r + tap * s]
                s = k // 2
                return wt.view(cin, cout, 2, s).permute(3, 1, 2, 0).reshape(s, cout, 2 * cin)
            out = torch.zeros(k, cout, 2 * cin, dtype=F32)                # kernel == stride: tap 1 is empty
            out[:, :, :cin] = wt.permute(2, 1, 0)
            return out

        emb = []
        for name, n in (("rvq_first", 1), ("rvq_rest", cfg.num_quantizers - 1)):
            for k in range(n):
                p = f"quantizer.{name}.vq.layers.{k}._codebook."
                emb.append(f(p + "embedding_sum") / f(p + "cluster_usage").clamp(min=cfg.codebook_eps)[:, None])
            w[f"{name}.proj"] = f(f"quantizer.{name}.output_proj.weight")[:, :, 0]
        w["codebooks"] = torch.stack(emb)
        w["pre_conv.w"], w["pre_conv.b"] = f("pre_conv.conv.weight").flatten(1), f("pre_conv.conv.bias")
        P = "pre_transformer."
        for k in ("input_proj", "output_proj"):
            w[f"t.{k}.w"], w[f"t.{k}.b"] = f(P + k + ".weight"), f(P + k + ".bias")
        w["t.norm"] = f(P + "norm.weight")
        for i in range(cfg.num_hidden_layers):
            L = f"{P}layers.{i}."
            w[f"t{i}.qkv"] = torch.cat([f(L + f"self_attn.{x}_proj.weight") for x in "qkv"], 0)
            w[f"t{i}.o"] = f(L + "self_attn.o_proj.weight")
            w[f"t{i}.gate"], w[f"t{i}.up"], w[f"t{i}.down"] = (f(L + f"mlp.{x}_proj.weight") for x in ("gate", "up", "down"))
            w[f"t{i}.n1"], w[f"t{i}.n2"] = f(L + "input_layernorm.weight"), f(L + "post_attention_layernorm.weight")
            w[f"t{i}.s1"], w[f"t{i}.s2"] = f(L + "self_attn_layer_scale.scale"), f(L + "mlp_layer_scale.scale")
        for j in range(len(cfg.upsampling_ratios)):
            w[f"u{j}.tr.w"] = packed_convtr(f(f"upsample.{j}.0.conv.weight"), False)
            w[f"u{j}.tr.b"] = f(f"upsample.{j}.0.conv.bias")
            q = f"upsample.{j}.1."
            w[f"u{j}.dw.w"], w[f"u{j}.dw.b"] = f(q + "dwconv.conv.weight")[:, 0, :], f(q + "dwconv.conv.bias")
            w[f"u{j}.ln.w"], w[f"u{j}.ln.b"] = f(q + "norm.weight"), f(q + "norm.bias")
            w[f"u{j}.pw1.w"], w[f"u{j}.pw1.b"] = f(q + "pwconv1.weight"), f(q + "pwconv1.bias")
            w[f"u{j}.pw2.w"], w[f"u{j}.pw2.b"] = f(q + "pwconv2.weight"), f(q + "pwconv2.bias")
            w[f"u{j}.gamma"] = f(q + "gamma")
        w["d.in.w"], w["d.in.b"] = f("decoder.0.conv.weight").flatten(1), f("decoder.0.conv.bias")
        for bi in range(len(cfg.upsample_rates)):
            p = f"decoder.{bi + 1}.block."
            w[f"d{bi}.a"], w[f"d{bi}.ib"] = snake(p + "0.")
            w[f"d{bi}.tr.w"], w[f"d{bi}.tr.b"] = packed_convtr(f(p + "1.conv.weight"), True), f(p + "1.conv.bias")
            for u in range(3):
                q = f"{p}{u + 2}."
                w[f"d{bi}.{u}.a1"], w[f"d{bi}.{u}.ib1"] = snake(q + "act1.")
                w[f"d{bi}.{u}.a2"], w[f"d{bi}.{u}.ib2"] = snake(q + "act2.")
                w[f"d{bi}.{u}.w1"], w[f"d{bi}.{u}.b1"] = f(q + "conv1.conv.weight").flatten(1), f(q + "conv1.conv.bias")
                w[f"d{bi}.{u}.w2"], w[f"d{bi}.{u}.b2"] = f(q + "conv2.conv.weight").flatten(1), f(q + "conv2.conv.bias")
        n = len(cfg.upsample_rates) + 1
        w["d.out.a"], w["d.out.ib"] = snake(f"decoder.{n}.")
        w["d.out.w"], w["d.out.b"] = f(f"decoder.{n + 1}.conv.weight").flatten(1), f(f"decoder.{n + 1}.conv.bias")
        self.w = {k: v.contiguous().to(dev) for k, v in w.items()}
        # tensor-core copies (tf32 hi/lo tiles) of the wide convolutions: key -> packed tiles
        convs = {"rvq_first.proj": 1, "rvq_rest.proj": 1, "pre_conv.w": 3, "t.input_proj.w": 1, "t.output_proj.w": 1, "d.in.w": 7,
                 "d.out.w": 7}
        for i in range(cfg.num_hidden_layers):
            convs.update({f"t{i}.{k}": 1 for k in ("qkv", "o", "gate", "up", "down")})
        for j in range(len(cfg.upsampling_ratios)):
            convs.update({f"u{j}.pw1.w": 1, f"u{j}.pw2.w": 1})
        for bi in range(len(cfg.upsample_rates)):
            for u in range(3):
                convs.update({f"d{bi}.{u}.w1": 7, f"d{bi}.{u}.w2": 1})
        self.tc: Dict[str, torch.Tensor] = {}
        for k, ks in convs.items():
            t = _tc.pack_conv(self.w[k], self.w[k].shape[1] // ks, ks, min_cin=96)
            if t is not None:
                self.tc[k] = t
        for k in [f"u{j}.tr.w" for j in range(len(cfg.upsampling_ratios))] + [f"d{bi}.tr.w" for bi in range(len(cfg.upsample_rates))]:
            t = _tc.pack_convtr(self.w[k], self.w[k].shape[2] // 2, min_cin=96)
            if t is not None:
                self.tc[k] = t
        return self

    # ---- cache --------------------------------------------------------------------------------------------
    def init_cache(self, batch_size: int, device=None, dtype: torch.dtype = F32,
                   detokenize_interval: int = 10) -> Qwen3TTSDecoderCache:
        """Zero state for ``batch_size`` streams (qwen3_codec.py:1381-1539 without the scratch buffers)."""
        cfg, dev = self.cfg, torch.device(device) if device is not None else self.device
        z = lambda *s: torch.zeros(*s, dtype=F32, device=dev)        # noqa: E731
        c = Qwen3TTSDecoderCache(
            attention_cache=z(batch_size, cfg.num_hidden_layers, cfg.num_key_value_heads, cfg.sliding_window, 2 * cfg.head_dim),
            position_offset=torch.zeros(batch_size, dtype=torch.long, device=dev), pre_conv_cache=z(batch_size, cfg.codebook_dim, 2),
            upsample_conv_caches=[z(batch_size, cfg.latent_dim, 6) for _ in cfg.upsampling_ratios],
            decoder_conv_caches=[z(batch_size, cfg.latent_dim, 6)], transconv_caches=[])
        ch = cfg.decoder_dim
        for _ in cfg.upsample_rates:
            c.transconv_caches.append(z(batch_size, ch, 1))
            ch //= 2
            c.decoder_conv_caches += [z(batch_size, ch, 6 * d) for d in (1, 3, 9)]
        c.decoder_conv_caches.append(z(batch_size, ch, 6))
        return c

    # ---- kernels ------------------------------------------------------------------------------------------
    def _conv(self, x, key, bias, B, cin, cout, T, ksize=1, dil=1, epi=0, resid=None, scale=None, ctx=None, act=0, a=None, ib=None):
        """``key``: name of the weight in self.w; the tcgen05 kernel takes the layer when a packed copy exists."""
        y = torch.empty(B, cout, T, dtype=F32, device=self.device)
        p = lambda t: None if t is None else t.data_ptr()            # noqa: E731
        tc = self.tc.get(key)
        if tc is not None and act:
            x, act, a, ib = self._activate(x, act, a, ib, B, cin, T), 0, None, None
        call("vb_codec_conv_tc" if tc is not None else "vb_codec_conv", y.data_ptr(), x.data_ptr(),
             (tc if tc is not None else self.w[key]).data_ptr(), p(bias), p(resid), p(scale), p(ctx), p(a), p(ib), epi, act, B, cin,
             cout, T, ksize, dil, ops._stream())
        return y

    def _convtr(self, x, key, bias, B, cin, cout, T, stride, ctx=None, act=0, a=None, ib=None):
        y = torch.empty(B, cout, T * stride, dtype=F32, device=self.device)
        p = lambda t: None if t is None else t.data_ptr()            # noqa: E731
        tc = self.tc.get(key)
        if tc is not None and act:
            x, act, a, ib = self._activate(x, act, a, ib, B, cin, T), 0, None, None
        call("vb_codec_convtr_tc" if tc is not None else "vb_codec_convtr", y.data_ptr(), x.data_ptr(),
             (tc if tc is not None else self.w[key]).data_ptr(), p(bias), p(ctx), p(a), p(ib), act, B, cin, cout, T, stride,
             ops._stream())
        return y

    def _activate(self, x, act, a, ib, B, C, T):
        """act(x) once, for a tensor-core convolution (its k taps would otherwise each re-evaluate it in the loaders)"""
        y = torch.empty_like(x)
        call("vb_codec_activate", y.data_ptr(), x.data_ptr(), a.data_ptr(), ib.data_ptr(), act, B, C, T, ops._stream())
        return y

    def _cache_update(self, cache, x, B, C, L, act=0, a=None, ib=None):
        p = lambda t: None if t is None else t.data_ptr()            # noqa: E731
        call("vb_codec_cache_update", cache.data_ptr(), x.data_ptr(), p(a), p(ib), act, B, C, L, cache.shape[2], ops._stream())

    @torch.no_grad()
    def decode_chunk(self, codes: torch.Tensor, decoder_cache: Optional[Qwen3TTSDecoderCache] = None):
        """codes [B, num_quantizers, T] -> (wav [B, 1, T * hop] fp32 in [-1, 1], the cache, updated in place)."""
        cfg, w, dev, st = self.cfg, self.w, self.device, ops._stream()
        if not codes.is_cuda:
            raise VoxB200Error("Qwen3TTSDecoder.decode_chunk needs CUDA tensors (there is no CPU path)")
        c = codes.to(torch.int64).contiguous()
        B, K, T = c.shape
        if K != cfg.num_quantizers:
            raise ValueError(f"Expected {cfg.num_quantizers} layer of codes, got {K}")
        if T >= cfg.sliding_window:
            raise VoxB200Error(f"chunks of {T} frames do not fit the {cfg.sliding_window}-slot attention window")
        cache = decoder_cache if decoder_cache is not None else self.init_cache(B)
        if cache.attention_cache.shape[0] != B:
            raise VoxB200Error("decoder cache and codes disagree on the batch size")
        Cb, Lt, Hd = cfg.codebook_dim, cfg.latent_dim, cfg.hidden_size
        e = dict(dtype=F32, device=dev)
        # ---- split RVQ decode ----
        dq = Cb // 2
        z = torch.empty(B, dq, T, **e)
        call("vb_mimi_codes_sum", z.data_ptr(), c.data_ptr(), w["codebooks"].data_ptr(), B, K, 0, 1, cfg.codebook_size, dq, T, st)
        h = self._conv(z, "rvq_first.proj", None, B, dq, Cb, T)
        z2 = torch.empty(B, dq, T, **e)
        call("vb_mimi_codes_sum", z2.data_ptr(), c.data_ptr(), w["codebooks"].data_ptr(), B, K, 1, K, cfg.codebook_size, dq, T, st)
        h = self._conv(z2, "rvq_rest.proj", None, B, dq, Cb, T, epi=1, resid=h)
        # ---- pre_conv (k3, cache 2) ----
        x = self._conv(h, "pre_conv.w", w["pre_conv.b"], B, Cb, Lt, T, ksize=3, ctx=cache.pre_conv_cache)
        self._cache_update(cache.pre_conv_cache, h, B, Cb, T)
        # ---- transformer ----
        H, Hkv, D, I = cfg.num_attention_heads, cfg.num_key_value_heads, cfg.head_dim, cfg.intermediate_size
        x = self._conv(x, "t.input_proj.w", w["t.input_proj.b"], B, Lt, Hd, T)
        for i in range(cfg.num_hidden_layers):
            n = torch.empty(B, Hd, T, **e)
            call("vb_codec_rmsnorm", n.data_ptr(), x.data_ptr(), w[f"t{i}.n1"].data_ptr(), B, Hd, T, cfg.rms_norm_eps, st)
            qkv = self._conv(n, f"t{i}.qkv", None, B, Hd, (H + 2 * Hkv) * D, T)
            a = torch.empty(B, H * D, T, **e)
            lc = cache.attention_cache[:, i]                # one layer of [B, layers, Hkv, W, 2 D]: items stride(0) apart
            call("vb_codec_attn_chunk", a.data_ptr(), qkv.data_ptr(), lc.data_ptr(), lc.stride(0), cache.position_offset.data_ptr(),
                 B, H, Hkv, D, T, cfg.sliding_window, float(cfg.rope_theta), st)
            x = self._conv(a, f"t{i}.o", None, B, H * D, Hd, T, epi=2, resid=x, scale=w[f"t{i}.s1"])
            call("vb_codec_rmsnorm", n.data_ptr(), x.data_ptr(), w[f"t{i}.n2"].data_ptr(), B, Hd, T, cfg.rms_norm_eps, st)
            g = self._conv(n, f"t{i}.gate", None, B, Hd, I, T, epi=4)
            m = self._conv(n, f"t{i}.up", None, B, Hd, I, T, epi=5, resid=g)
            x = self._conv(m, f"t{i}.down", None, B, I, Hd, T, epi=2, resid=x, scale=w[f"t{i}.s2"])
        cache.position_offset.add_(T)
        n = torch.empty(B, Hd, T, **e)
        call("vb_codec_rmsnorm", n.data_ptr(), x.data_ptr(), w["t.norm"].data_ptr(), B, Hd, T, cfg.rms_norm_eps, st)
        x = self._conv(n, "t.output_proj.w", w["t.output_proj.b"], B, Hd, Lt, T)
        # ---- upsampling: ConvTranspose (kernel == stride) + ConvNeXt ----
        for j, f in enumerate(cfg.upsampling_ratios):
            x, T = self._convtr(x, f"u{j}.tr.w", w[f"u{j}.tr.b"], B, Lt, Lt, T, f), T * f
            uc = cache.upsample_conv_caches[j]
            d = torch.empty(B, Lt, T, **e)
            call("vb_codec_dwconv", d.data_ptr(), x.data_ptr(), w[f"u{j}.dw.w"].data_ptr(), w[f"u{j}.dw.b"].data_ptr(), uc.data_ptr(),
                 B, Lt, T, 7, st)
            self._cache_update(uc, x, B, Lt, T)
            ln = torch.empty(B, Lt, T, **e)
            call("vb_mimi_layernorm", ln.data_ptr(), d.data_ptr(), w[f"u{j}.ln.w"].data_ptr(), w[f"u{j}.ln.b"].data_ptr(), B, Lt, T,
                 1e-6, st)
            m = self._conv(ln, f"u{j}.pw1.w", w[f"u{j}.pw1.b"], B, Lt, 4 * Lt, T, epi=3)
            x = self._conv(m, f"u{j}.pw2.w", w[f"u{j}.pw2.b"], B, 4 * Lt, Lt, T, epi=2, resid=x, scale=w[f"u{j}.gamma"])
        # ---- decoder ----
        dc, ci, ch = cache.decoder_conv_caches, 0, cfg.decoder_dim
        y = self._conv(x, "d.in.w", w["d.in.b"], B, Lt, ch, T, ksize=7, ctx=dc[ci])
        self._cache_update(dc[ci], x, B, Lt, T)
        x, ci = y, ci + 1
        for bi, rate in enumerate(cfg.upsample_rates):
            a_, ib_ = w[f"d{bi}.a"], w[f"d{bi}.ib"]
            tc = cache.transconv_caches[bi]
            y = self._convtr(x, f"d{bi}.tr.w", w[f"d{bi}.tr.b"], B, ch, ch // 2, T, rate, ctx=tc, act=2, a=a_, ib=ib_)
            self._cache_update(tc, x, B, ch, T, act=2, a=a_, ib=ib_)
            x, ch, T = y, ch // 2, T * rate
            for u, dil in enumerate((1, 3, 9)):
                a1, ib1, a2, ib2 = (w[f"d{bi}.{u}.{k}"] for k in ("a1", "ib1", "a2", "ib2"))
                r = self._conv(x, f"d{bi}.{u}.w1", w[f"d{bi}.{u}.b1"], B, ch, ch, T, ksize=7, dil=dil, ctx=dc[ci], act=2, a=a1,
                               ib=ib1)
                self._cache_update(dc[ci], x, B, ch, T, act=2, a=a1, ib=ib1)
                ci += 1
                x = self._conv(r, f"d{bi}.{u}.w2", w[f"d{bi}.{u}.b2"], B, ch, ch, T, epi=1, resid=x, act=2, a=a2, ib=ib2)
        wav = self._conv(x, "d.out.w", w["d.out.b"], B, ch, 1, T, ksize=7, epi=6, ctx=dc[ci], act=2, a=w["d.out.a"],
                         ib=w["d.out.ib"])
        self._cache_update(dc[ci], x, B, ch, T, act=2, a=w["d.out.a"], ib=w["d.out.ib"])
        return wav, cache

    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """Non-streaming convenience: one chunk from a zero cache."""
        return self.decode_chunk(codes, None)[0]
