"""SNAC decode path on the sm_100a kernels (drop-in for ``vox_serve/tokenizer/snac.py`` decode:
``SNAC.from_config`` / ``from_pretrained`` / ``load_state_dict`` / ``decode(codes)``, snac.py:438-466).

Only the decoder + RVQ ``from_codes`` exist here (the encoder is not on the serving path).  At load time
the weight-norm parametrisation is folded once (the reference re-derives ``g*v/||v||`` on every forward,
snac.py:244-249), transposed-conv weights are repacked per output phase, and every Snake is assigned to
the epilogue of the stage that produces its input.  ``NoiseBlock`` noise (``torch.randn`` in the reference,
snac.py:208) comes from the library's own Philox normal generator (``vb_randn``: one launch for the four blocks,
device-side offset, CUDA-graph replayable) unless explicit noise tensors / a ``noise_source`` are given.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from .. import ops
from .. import _lib
from .._lib import VoxB200Error, call


def _fold(sd, prefix: str) -> torch.Tensor:
    g = sd[prefix + ".parametrizations.weight.original0"].float()
    v = sd[prefix + ".parametrizations.weight.original1"].float()
    n = torch.linalg.vector_norm(v, ord=2, dim=tuple(range(1, v.dim())), keepdim=True)
    return v * (g / n)


class SNAC:
    def __init__(self, sampling_rate=24000, encoder_dim=48, encoder_rates=(2, 4, 8, 8), latent_dim=None,
                 decoder_dim=1024, decoder_rates=(8, 8, 4, 2), attn_window_size=None, codebook_size=4096,
                 codebook_dim=8, vq_strides=(4, 2, 1), noise=True, depthwise=True, enable_torch_compile=False,
                 device="cuda", **_):
        if attn_window_size is not None:
            raise VoxB200Error("SNAC variants with LocalMHA (attn_window_size) are not on the Orpheus path")
        if not depthwise or not noise or len(vq_strides) != 3:
            raise VoxB200Error("only the depthwise + noise, 3-codebook SNAC configuration (snac_24khz) is built")
        self.sampling_rate = sampling_rate
        self.latent_dim = latent_dim or encoder_dim * (2 ** len(encoder_rates))
        self.decoder_dim, self.decoder_rates = decoder_dim, tuple(decoder_rates)
        self.codebook_size, self.codebook_dim, self.vq_strides = codebook_size, codebook_dim, tuple(vq_strides)
        self.hop_length = math.prod(encoder_rates)
        self.device = torch.device(device)
        self.w: Dict[str, torch.Tensor] = {}
        self.loaded = False
        # NoiseBlock input (snac.py:206-212): None = torch.randn on the current CUDA generator like the reference;
        # a callable(shapes) -> tensors lets tests inject the oracle's noise
        self.noise_source = None
        self.noise_state = None          # device {seed, offset, arrivals} of the default noise stream (ops.randn)

    # ---- loading -----------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config_path, enable_torch_compile=False, device="cuda"):
        with open(config_path) as f:
            cfg = json.load(f)
        return cls(**cfg, device=device)

    @classmethod
    def from_pretrained(cls, repo_id, enable_torch_compile=False, device="cuda", **kw):
        if not os.path.isdir(repo_id):
            raise VoxB200Error("no network in this build: pass a local directory with config.json + pytorch_model.bin")
        m = cls.from_config(os.path.join(repo_id, "config.json"), device=device)
        m.load_state_dict(torch.load(os.path.join(repo_id, "pytorch_model.bin"), map_location="cpu"))
        return m

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device and self.loaded:
            self.w = {k: v.to(device) for k, v in self.w.items()}
            self.device = torch.device(device)
            self._pack_tensor_core_weights()
        self.device = torch.device(device)
        return self

    def load_state_dict(self, sd, strict: bool = False):
        dev, w = self.device, {}
        C, nq = self.latent_dim, len(self.vq_strides)
        w["codebooks"] = torch.stack([sd[f"quantizer.quantizers.{i}.codebook.weight"].float() for i in range(nq)])
        w["proj_w"] = torch.stack([_fold(sd, f"quantizer.quantizers.{i}.out_proj")[:, :, 0] for i in range(nq)])
        w["proj_b"] = torch.stack([sd[f"quantizer.quantizers.{i}.out_proj.bias"].float() for i in range(nq)])
        w["in_dw_w"] = _fold(sd, "decoder.model.0")[:, 0, :]
        w["in_dw_b"] = sd["decoder.model.0.bias"].float()
        w["in_pw_w"] = _fold(sd, "decoder.model.1")[:, :, 0]
        w["in_pw_b"] = sd["decoder.model.1.bias"].float()
        li = 2
        for bi, s in enumerate(self.decoder_rates):
            p = f"decoder.model.{li}"
            w[f"b{bi}.alpha"] = sd[p + ".block.0.alpha"].float().reshape(-1)
            wt = _fold(sd, p + ".block.1")                       # [Cin, Cout, 2s]
            cin, cout, k = wt.shape
            assert k == 2 * s
            # packed[r][co][tap*Cin + ci] = W[ci][co][r + tap*s]
            w[f"b{bi}.ct_w"] = wt.view(cin, cout, 2, s).permute(3, 1, 2, 0).reshape(s, cout, 2 * cin)
            w[f"b{bi}.ct_b"] = sd[p + ".block.1.bias"].float()
            w[f"b{bi}.noise_w"] = _fold(sd, p + ".block.2.linear")[:, :, 0]
            for j in range(3):
                q = f"{p}.block.{3 + j}"
                w[f"b{bi}.r{j}.alpha1"] = sd[q + ".block.0.alpha"].float().reshape(-1)
                w[f"b{bi}.r{j}.dw_w"] = _fold(sd, q + ".block.1")[:, 0, :]
                w[f"b{bi}.r{j}.dw_b"] = sd[q + ".block.1.bias"].float()
                w[f"b{bi}.r{j}.alpha2"] = sd[q + ".block.2.alpha"].float().reshape(-1)
                w[f"b{bi}.r{j}.pw_w"] = _fold(sd, q + ".block.3")[:, :, 0]
                w[f"b{bi}.r{j}.pw_b"] = sd[q + ".block.3.bias"].float()
            li += 1
        w["out.alpha"] = sd[f"decoder.model.{li}.alpha"].float().reshape(-1)
        w["out.w"] = _fold(sd, f"decoder.model.{li + 1}")[0]     # [C, 7]
        w["out.b"] = sd[f"decoder.model.{li + 1}.bias"].float()
        self.w = {k: v.contiguous().to(dev) for k, v in w.items()}
        self._pack_tensor_core_weights()
        self.loaded = True
        self.noise_state = None
        return self

    def ensure_noise_state(self) -> torch.Tensor:
        """Create the device {seed, offset, arrivals} of the default noise stream now (a host -> device copy: must not
        happen inside a CUDA-graph capture)."""
        if self.noise_state is None:
            dev = self.device
            seed = torch.cuda.default_generators[dev.index or 0].initial_seed() & ((1 << 62) - 1)
            self.noise_state = torch.tensor([seed ^ 0x534E4143, 0, 0], dtype=torch.int64, device=dev)
        return self.noise_state

    def _pack_tensor_core_weights(self):
        """The 1x1 and transposed-conv weights whose input width is a multiple of 32 channels are re-tiled for the
        tcgen05 tf32 hi/lo kernel (vb_snac_pack_tf32x3); narrower layers (test-sized configs) stay on the SIMT
        kernels.  VB_SNAC_FP32=1 keeps everything on the SIMT kernels."""
        self.tc: Dict[str, torch.Tensor] = {}
        if self.device.type != "cuda" or os.environ.get("VB_SNAC_FP32", "0") == "1":
            return
        st = ops._stream()
        for k, v in self.w.items():
            if not (k.endswith("pw_w") or k.endswith("noise_w") or k.endswith("ct_w")):
                continue
            phases, M, K = (1, *v.shape) if v.dim() == 2 else tuple(v.shape)
            n = _lib.load().vb_snac_tf32x3_bytes(phases, M, K)
            if n <= 0 or K < int(os.environ.get("VB_SNAC_TC_MIN_K", "128")):
                continue
            dst = torch.empty(n, dtype=torch.uint8, device=self.device)
            call("vb_snac_pack_tf32x3", dst.data_ptr(), v.data_ptr(), phases, M, K, st)
            self.tc[k] = dst

    def _pwconv(self, key, y, x, bias, resid, noise, alpha, epi, B, cin, cout, T, lo, hi, st):
        tc = self.tc.get(key)
        call("vb_snac_pwconv_tc" if tc is not None else "vb_snac_pwconv", y.data_ptr(), x.data_ptr(),
             (tc if tc is not None else self.w[key]).data_ptr(), None if bias is None else bias.data_ptr(),
             None if resid is None else resid.data_ptr(), None if noise is None else noise.data_ptr(),
             None if alpha is None else alpha.data_ptr(), epi, B, cin, cout, T, lo, hi, st)

    def _convtr(self, key, y, x, bias, B, cin, cout, T, s, lo, hi, st):
        tc = self.tc.get(key)
        call("vb_snac_convtr_tc" if tc is not None else "vb_snac_convtr", y.data_ptr(), x.data_ptr(),
             (tc if tc is not None else self.w[key]).data_ptr(), bias.data_ptr(), None, B, cin, cout, T, s, lo, hi, st)

    def synthetic_state_dict(self, seed: int = 0) -> Dict[str, torch.Tensor]:
        """Seeded weights under the reference's state_dict names (weight-norm ``original0`` = g, ``original1`` = v;
        snac.py:244-249) for runs without the HF checkpoint (no network).  Scales keep activations O(1)."""
        gen = torch.Generator().manual_seed(seed)
        sd: Dict[str, torch.Tensor] = {}

        def wn(prefix, shape, fan_in, bias_dim=None, gain=1.0):
            v = torch.randn(*shape, generator=gen)
            rowsize = math.prod(shape[1:])
            jitter = 1.0 + 0.2 * (2 * torch.rand(shape[0], *([1] * (len(shape) - 1)), generator=gen) - 1)
            sd[prefix + ".parametrizations.weight.original0"] = gain * math.sqrt(rowsize / fan_in) * jitter
            sd[prefix + ".parametrizations.weight.original1"] = v
            if bias_dim is not None:
                sd[prefix + ".bias"] = torch.randn(bias_dim, generator=gen) * 0.05

        L, D, nq = self.latent_dim, self.decoder_dim, len(self.vq_strides)
        for i in range(nq):
            p = f"quantizer.quantizers.{i}"
            sd[p + ".codebook.weight"] = torch.randn(self.codebook_size, self.codebook_dim, generator=gen)
            wn(p + ".out_proj", (L, self.codebook_dim, 1), self.codebook_dim * nq, L)
        wn("decoder.model.0", (L, 1, 7), 7, L)
        wn("decoder.model.1", (D, L, 1), L, D)
        li, cin = 2, D
        for s in self.decoder_rates:
            cout, p = cin // 2, f"decoder.model.{li}"
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

    # ---- decode ------------------------------------------------------------------------------
    def noise_shapes(self, batch: int, t_latent: int):
        out, t = [], t_latent
        for s in self.decoder_rates:
            t *= s
            out.append((batch, 1, t))
        return out

    def _stage_ranges(self, t_latent: int, out_range):
        """Receptive field of the kept output slice, walked backwards through the decoder: for every DecoderBlock
        the output ranges its stages must cover (all half-open, clamped).  Returns a list, first block first, of
        (convtr_out, [unit0, unit1, unit2]) plus the final conv's input range."""
        rates = self.decoder_rates
        lens = [t_latent]
        for s in rates:
            lens.append(lens[-1] * s)
        lo, hi = out_range
        lo, hi = max(0, lo - 3), min(lens[-1], hi + 3)          # final conv k7, padding 3
        blocks = [None] * len(rates)
        for b in reversed(range(len(rates))):
            T = lens[b + 1]
            units = [None] * 3
            cur = (lo, hi)
            for j, dil in reversed(list(enumerate((1, 3, 9)))):
                units[j] = cur                                    # unit j output (and its depthwise conv) range
                cur = (max(0, cur[0] - 3 * dil), min(T, cur[1] + 3 * dil))
            blocks[b] = (cur, units)                              # cur: convtr / NoiseBlock output range
            s = rates[b]
            pad = (s + 1) // 2
            lo = max(0, (cur[0] + pad - (2 * s - 1)) // s)        # ConvTranspose1d taps: to = ti*s - pad + k
            hi = min(lens[b], (cur[1] - 1 + pad) // s + 1)
        return blocks, (lo, hi)

    @torch.no_grad()
    def decode(self, codes: List[torch.Tensor], noises: Optional[Sequence[torch.Tensor]] = None,
               out_range: Optional[Sequence[int]] = None) -> torch.Tensor:
        """codes: 3 int tensors [B, T/stride_i] -> waveform [B, 1, T*prod(rates)] fp32 (snac.py:438-441).
        out_range=(t0, t1) returns only those samples (what OrpheusModel.postprocess keeps, orpheus.py:506) and
        computes only their receptive field in every stage -- bit-identical to slicing the full decode."""
        if not self.loaded:
            raise VoxB200Error("SNAC weights not loaded")
        w, st = self.w, ops._stream()
        dev = self.device
        c = [x.to(device=dev, dtype=torch.int32).contiguous() for x in codes]
        B = c[0].shape[0]
        T = c[-1].shape[1] * self.vq_strides[-1]
        C = self.latent_dim
        T_out = T * math.prod(self.decoder_rates)
        t0, t1 = (0, T_out) if out_range is None else (int(out_range[0]), int(out_range[1]))
        blocks, (in_lo, in_hi) = self._stage_ranges(T, (t0, t1))
        f32 = dict(dtype=torch.float32, device=dev)
        z = torch.empty(B, C, T, **f32)
        call("vb_snac_from_codes", z.data_ptr(), c[0].data_ptr(), c[1].data_ptr(), c[2].data_ptr(),
             w["codebooks"].data_ptr(), w["proj_w"].data_ptr(), w["proj_b"].data_ptr(), B, C, T,
             self.codebook_size, self.codebook_dim, *self.vq_strides, st)
        x = torch.empty_like(z)
        call("vb_snac_dwconv7", x.data_ptr(), z.data_ptr(), w["in_dw_w"].data_ptr(), w["in_dw_b"].data_ptr(), None,
             None, B, C, T, 1, in_lo, in_hi, st)
        ch = self.decoder_dim
        y = torch.empty(B, ch, T, **f32)
        # 1x1 conv 768 -> 1024; its output only feeds block 0's Snake -> fuse that Snake here
        self._pwconv("in_pw_w", y, x, w["in_pw_b"], None, None, w["b0.alpha"], 0, B, C, ch, T, in_lo, in_hi, st)
        x = y
        nb = len(self.decoder_rates)
        if noises is None:
            shapes = self.noise_shapes(B, T)
            if self.noise_source is not None:
                noises = self.noise_source(shapes)
            else:
                # one launch for the four NoiseBlock inputs (the reference draws torch.randn per block, snac.py:208);
                # the device-side offset advances per call, so a captured vocoder graph draws fresh noise every replay
                sizes = [math.prod(s) for s in shapes]
                flat = ops.randn(torch.empty(sum(sizes), **f32), rng_state=self.ensure_noise_state())
                noises = [t.view(s) for t, s in zip(torch.split(flat, sizes), shapes)]
        for bi, s in enumerate(self.decoder_rates):
            (c_lo, c_hi), units = blocks[bi]
            cin, cout = ch, ch // 2
            u = torch.empty(B, cout, T * s, **f32)
            self._convtr(f"b{bi}.ct_w", u, x, w[f"b{bi}.ct_b"], B, cin, cout, T, s, c_lo, c_hi, st)
            T *= s
            nz = noises[bi].to(device=dev, dtype=torch.float32).contiguous()
            assert nz.numel() == B * T, "noise tensor shape mismatch"
            x = torch.empty_like(u)
            self._pwconv(f"b{bi}.noise_w", x, u, None, None, nz, None, 2, B, cout, cout, T, c_lo, c_hi, st)
            for j, dil in enumerate((1, 3, 9)):
                u_lo, u_hi = units[j]
                h = torch.empty_like(x)
                call("vb_snac_dwconv7", h.data_ptr(), x.data_ptr(), w[f"b{bi}.r{j}.dw_w"].data_ptr(),
                     w[f"b{bi}.r{j}.dw_b"].data_ptr(), w[f"b{bi}.r{j}.alpha1"].data_ptr(),
                     w[f"b{bi}.r{j}.alpha2"].data_ptr(), B, cout, T, dil, u_lo, u_hi, st)
                # the last unit's output only feeds the next stage's Snake: fuse it into this epilogue
                nxt = None
                if j == 2:
                    nxt = w[f"b{bi + 1}.alpha"] if bi + 1 < nb else w["out.alpha"]
                o = torch.empty_like(x)
                self._pwconv(f"b{bi}.r{j}.pw_w", o, h, w[f"b{bi}.r{j}.pw_b"], x, None, nxt, 1, B, cout, cout, T,
                             u_lo, u_hi, st)
                x = o
            ch = cout
        wav = torch.empty(B, 1, t1 - t0, **f32)
        call("vb_snac_final", wav.data_ptr(), x.data_ptr(), w["out.w"].data_ptr(), w["out.b"].data_ptr(), None, B,
             ch, T, t0, t1, st)
        return wav
