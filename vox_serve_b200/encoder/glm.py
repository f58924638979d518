"""GLM-4-Voice speech tokenizer on the B200 kernels: the Whisper-style VQ encoder that turns the log-mel features of
a spoken prompt into ``<|audio_N|>`` ids (``vox_serve/encoder/glm.py:84-369``; the STS prompt side, SURVEY.md §8f
row 3).  Same class names, constructor arguments and call signatures as the reference module:
``GLMWhisperVQEncoder(config)(input_features, attention_mask) -> ids`` and ``GLMVoiceEncoder(...).encode(audio)``.

How it maps to the kernels (everything bf16, token-major ``[T][C]``, rounded where the reference's bf16 modules round):

* the two causal k = 3 convolutions (``CausalConv1d``, :84-107) are GEMMs whose activation operand is a tensor map
  over OVERLAPPING rows of the left-padded, token-major input -- row t of the operand is the 3 C consecutive values
  of input rows t .. t + 2 (stride 1) or 2t .. 2t + 2 (stride 2): no im2col copy (``vb_chw_to_rows`` lays the
  channels-first mel features out once);
* q / k / v / out / fc1 / fc2: ``gemm_bf16_kernel`` with its bias epilogue; K and V are written by their projections
  straight into a one-page "paged" buffer, so the attention kernel needs no append step;
* attention: ``paged_prefill_attn_kernel`` (20 kv heads, no grouping, head_dim 64) with a PER-ROW key bound -- the
  reference's ``(causal | same 200-frame block) & key-not-padding`` mask (:260-277) lets row i see exactly the keys
  below ``min(end of i's block, valid length)``;
* LayerNorm (+ the residual add before it), erf-GELU (+ the position add), average pooling and the codebook arg-min
  (``vector_quantize`` :247-258, on fp32 ``x c^T`` from a mode-1 GEMM): ``csrc/encoder.cu``.

There is no CPU path: the kernels raise on anything but CUDA tensors.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch

from .. import ops
from .._lib import VoxB200Error, call

BF16 = torch.bfloat16


@dataclass
class GLMEncoderConfig:
    """The fields of the reference's ``GLMEncoderConfig`` (glm.py:14-81) that shape the computation (the others are
    training-time / decoder-side settings of the Whisper checkpoint's config.json and are accepted and ignored)."""
    d_model: int = 1280
    encoder_attention_heads: int = 20
    encoder_ffn_dim: int = 5120
    encoder_layers: int = 32
    num_mel_bins: int = 128
    max_source_positions: int = 1500
    pooling_kernel_size: int = 4
    pooling_position: int = 16
    quantize_position: int = 16
    quantize_vocab_size: int = 16384
    quantize_causal_block_size: int = 200
    extra: Dict[str, Any] = field(default_factory=dict)

    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]) -> "GLMEncoderConfig":
        names = {f for f in cls.__dataclass_fields__ if f != "extra"}
        return cls(**{k: v for k, v in config_dict.items() if k in names},
                   extra={k: v for k, v in config_dict.items() if k not in names})


def _stream():
    return torch.cuda.current_stream().cuda_stream


class GLMWhisperVQEncoder:
    """``GLMWhisperVQEncoder`` (glm.py:217-323).  ``state_dict``: the reference module's parameter names."""

    def __init__(self, config: GLMEncoderConfig, state_dict: Optional[Dict[str, torch.Tensor]] = None, device="cuda"):
        if not torch.cuda.is_available():
            raise VoxB200Error("GLMWhisperVQEncoder needs a CUDA device: there is no CPU path")
        self.config = config
        self.device = torch.device(device)
        self.d_model, self.n_heads = config.d_model, config.encoder_attention_heads
        self.head_dim = self.d_model // self.n_heads
        if self.head_dim not in (64, 128) or self.d_model % 64 != 0:
            raise VoxB200Error(f"head_dim {self.head_dim} unsupported by the attention kernel (64, 128)")
        self.n_layers = config.quantize_position
        self.loaded = False
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ---- weights ------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        dev, D = self.device, self.d_model

        def w(name):
            if name not in sd:
                raise VoxB200Error(f"GLMWhisperVQEncoder: missing weight '{name}'")
            return sd[name].to(device=dev, dtype=BF16).contiguous()

        def packed(t):
            return ops.pack_weight(t.contiguous(), 128)

        # Conv1d weight [out, in, k] -> GEMM weight [out, (k, in)]: tap-major, matching the overlapping-row operand
        self.conv1_w = packed(w("conv1.weight").permute(0, 2, 1).reshape(D, -1))
        self.conv2_w = packed(w("conv2.weight").permute(0, 2, 1).reshape(D, -1))
        self.conv1_b, self.conv2_b = w("conv1.bias"), w("conv2.bias")
        self.embed_positions = w("embed_positions.weight")
        cb = w("codebook.weight")
        self.codebook = packed(cb)
        # load-time constant of vector_quantize (glm.py:250), computed exactly as the reference computes it per call
        self.codebook_sqr = torch.sum(cb ** 2, dim=1).contiguous()
        self.layers: List[Dict[str, Any]] = []
        for i in range(self.n_layers):
            p = f"layers.{i}."
            L = {}
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                L[n] = packed(w(p + f"self_attn.{n}.weight"))
                L[n + "_b"] = w(p + f"self_attn.{n}.bias") if n != "k_proj" else None
            L["fc1"], L["fc1_b"] = packed(w(p + "fc1.weight")), w(p + "fc1.bias")
            L["fc2"], L["fc2_b"] = packed(w(p + "fc2.weight")), w(p + "fc2.bias")
            for n in ("self_attn_layer_norm", "final_layer_norm"):
                L[n + "_w"], L[n + "_b"] = w(p + n + ".weight"), w(p + n + ".bias")
            self.layers.append(L)
        self.loaded = True
        return self

    def to(self, *args, **kwargs):           # the reference calls .to(dtype).to(device) on the module (glm.py:342)
        return self

    def eval(self):
        return self

    # ---- pieces ---------------------------------------------------------------------------------------
    def _add_layernorm(self, h, delta, wname, L, y: Optional[ops.TiledAct] = None):
        """h += delta (in place); y = LayerNorm(h) in the tiled layout the next projection streams (None: add only)."""
        call("vb_add_layernorm", None if y is None else y.data.data_ptr(), h.data_ptr(),
             None if delta is None else delta.data_ptr(), None if y is None else L[wname + "_w"].data_ptr(),
             None if y is None else L[wname + "_b"].data_ptr(), h.shape[0], h.shape[1], 1e-5,
             0 if y is None else y.t_tile, _stream())
        return y

    def _gelu(self, x, out=None, add=None):
        """out = gelu(x) (+ add); ``out`` a row-major tensor (default: in place) or a TiledAct."""
        out = x if out is None else out
        tiled = isinstance(out, ops.TiledAct)
        call("vb_gelu_add", out.data.data_ptr() if tiled else out.data_ptr(), x.data_ptr(),
             None if add is None else add.data_ptr(), x.numel(), x.shape[-1], out.t_tile if tiled else 0, _stream())
        return out

    @staticmethod
    def block_causal_bounds(valid: int, T: int, block: int, device) -> torch.Tensor:
        """Keys visible to each row under the reference's mask (glm.py:260-277) with padding at the end only."""
        i = torch.arange(T, dtype=torch.int32)
        return torch.minimum((i // block + 1) * block, torch.tensor(valid, dtype=torch.int32)).to(device)

    def _attention_plan(self, T: int, bounds: torch.Tensor):
        page = (T + 31) // 32 * 32
        plan = ops.RowPlan(max(T, 8), self.device)
        plan.row_kvlen[:T] = bounds
        i32 = dict(dtype=torch.int32, device=self.device)
        plan.qo_indptr = torch.tensor([0, T], **i32)
        plan.kv_indptr = torch.tensor([0, 1], **i32)
        plan.kv_indices = torch.tensor([0], **i32)
        plan.n_req, plan.n_rows = 1, T
        kv = torch.zeros(1, 1, 2, page, self.n_heads, self.head_dim, dtype=BF16, device=self.device)
        return plan, kv, page

    def _layer(self, h, y, L, plan, kv, page, T, buf):
        """One encoder layer after its first LayerNorm (y); returns the MLP output the caller adds to h.  Every
        projection reads its activations in the tiled layout (one linear bulk copy per pipeline stage; a row-major
        operand costs the TMA unit one request per token row, 256 per stage at these widths)."""
        D, H, hd = self.d_model, self.n_heads, self.head_dim
        q = ops.gemm(y, L["q_proj"], bias=L["q_proj_b"])
        ops.gemm(y, L["k_proj"], out=kv[0, 0, 0].view(page, D)[:T])
        ops.gemm(y, L["v_proj"], out=kv[0, 0, 1].view(page, D)[:T], bias=L["v_proj_b"])
        a = ops.paged_attn(q.view(T, H, hd), kv, 0, plan, T, H, page, 0, None, out=buf["a"], prefill_tiles=True)
        o = ops.gemm(a, L["out_proj"], bias=L["out_proj_b"])
        y2 = self._add_layernorm(h, o, "final_layer_norm", L, y=buf["y"])     # h += attn ; y2 = LN2(h)
        f = ops.gemm(y2, L["fc1"], bias=L["fc1_b"])
        g = self._gelu(f, out=buf["f"])
        return ops.gemm(g, L["fc2"], bias=L["fc2_b"])                         # the caller adds it to h

    def _buffers(self, T: int):
        D, F = self.d_model, self.config.encoder_ffn_dim
        return {"y": ops.TiledAct(T, D, self.device), "a": ops.TiledAct(T, D, self.device),
                "f": ops.TiledAct(T, F, self.device)}

    # ---- forward (glm.py:279-323) ---------------------------------------------------------------------
    def forward(self, input_features: torch.Tensor, attention_mask: torch.Tensor, return_states: bool = False):
        if not self.loaded:
            raise VoxB200Error("GLMWhisperVQEncoder: no weights loaded")
        if not input_features.is_cuda:
            raise VoxB200Error("GLMWhisperVQEncoder runs on CUDA tensors only")
        cfg, D = self.config, self.d_model
        B, M, frames = input_features.shape
        if frames % 2 != 0:
            raise VoxB200Error("feature frames must be even (the extractor pads to a multiple of the encoder stride)")
        out_ids, states = [], []
        for b in range(B):
            feats = input_features[b].to(BF16).contiguous()
            valid = int(attention_mask[b, ::2].sum())
            T = frames // 2
            if T > self.embed_positions.shape[0]:
                raise VoxB200Error(f"{T} positions exceed max_source_positions {self.embed_positions.shape[0]}")
            # conv1 + GELU: rows [2 + frames][M] -> [frames][D], written behind two zero rows for conv2's padding
            x0 = torch.empty(2 + frames, M, dtype=BF16, device=self.device)
            call("vb_chw_to_rows", x0.data_ptr(), feats.data_ptr(), M, frames, 2, _stream())
            c1 = ops.gemm(x0.as_strided((frames, 3 * M), (M, 1)), self.conv1_w, bias=self.conv1_b)
            x1 = torch.zeros(2 + frames, D, dtype=BF16, device=self.device)
            self._gelu(c1, out=x1[2:])
            # conv2 (stride 2) + GELU + positions
            c2 = ops.gemm(x1.as_strided((T, 3 * D), (2 * D, 1)), self.conv2_w, bias=self.conv2_b)
            h = self._gelu(c2, add=self.embed_positions[:T].contiguous())
            block = cfg.quantize_causal_block_size
            plan, kv, page = self._attention_plan(T, self.block_causal_bounds(valid, T, block, self.device))
            buf = self._buffers(T)
            ids = hidden_last = pooled = None
            delta = None
            for i, L in enumerate(self.layers):
                y = self._add_layernorm(h, delta, "self_attn_layer_norm", L, y=buf["y"])  # h += previous MLP ; y = LN1(h)
                delta = self._layer(h, y, L, plan, kv, page, T, buf)
                if i + 1 == cfg.pooling_position and cfg.pooling_kernel_size is not None:
                    self._add_layernorm(h, delta, "", L)
                    delta = None
                    hidden_last = h
                    k = cfg.pooling_kernel_size
                    Tp = (T + k - 1) // k
                    hp = torch.empty(Tp, D, dtype=BF16, device=self.device)
                    call("vb_avgpool_rows", hp.data_ptr(), h.data_ptr(), T, D, k, _stream())
                    h, T = hp, Tp
                    valid = int(attention_mask[b, ::2][::k].sum())
                    plan, kv, page = self._attention_plan(T, self.block_causal_bounds(valid, T, block // k, self.device))
                    buf = self._buffers(T)
                if i + 1 == cfg.quantize_position and cfg.quantize_vocab_size is not None:
                    if delta is not None:
                        self._add_layernorm(h, delta, "", L)
                        delta = None
                    pooled = h
                    acc = ops.gemm(h, self.codebook, mode=1)                  # fp32 [1][T][vocab] = x c^T
                    ids = torch.empty(T, dtype=torch.int64, device=self.device)
                    call("vb_vq_argmin", ids.data_ptr(), acc.data_ptr(), h.data_ptr(), self.codebook_sqr.data_ptr(), T,
                         self.codebook.N, D, _stream())
                    break
            out_ids.append(ids)
            states.append((hidden_last, pooled))
        ids = torch.stack(out_ids, 0)
        return (ids, states) if return_states else ids

    __call__ = forward


class GLMVoiceEncoder:
    """``GLMVoiceEncoder`` (glm.py:326-369): feature extractor + encoder.  The reference downloads the checkpoint
    (``hf_hub_download``); here ``repo_id`` is a local directory, or ``config`` / ``state_dict`` / ``feature_extractor``
    are handed in (how the tests and the offline GPU box build it)."""

    def __init__(self, repo_id: str = "THUDM/glm-4-voice-tokenizer", dtype: torch.dtype = BF16, device: str = "cuda",
                 config: Optional[GLMEncoderConfig] = None, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 feature_extractor=None):
        if dtype != BF16:
            raise VoxB200Error("the B200 encoder computes in bf16 only")
        self.repo_id, self.device, self.dtype = repo_id, device, dtype
        if config is None or state_dict is None:
            # a local checkpoint directory (config.json + *.safetensors), like the LM adapters (model/orpheus.py)
            import glob
            import json
            import os

            from safetensors.torch import load_file

            if not os.path.isdir(repo_id):
                raise VoxB200Error(f"'{repo_id}' is not a local checkpoint directory: pass config= and state_dict=, or "
                                   "download the tokenizer checkpoint first (the serving box has no network)")
            with open(os.path.join(repo_id, "config.json")) as f:
                config = GLMEncoderConfig.from_dict(json.load(f))
            state_dict = {}
            for fn in sorted(glob.glob(os.path.join(repo_id, "*.safetensors"))):
                state_dict.update(load_file(fn))
        self.config = config
        self.encoder = GLMWhisperVQEncoder(config, state_dict, device=device)
        if feature_extractor is None:
            from transformers import WhisperFeatureExtractor

            feature_extractor = WhisperFeatureExtractor.from_pretrained(repo_id)
        self.feature_extractor = feature_extractor
        k = config.pooling_kernel_size or 1
        self.stride = 1 * 2 * k * self.feature_extractor.hop_length          # conv strides x pooling x hop

    def encode(self, audio: torch.Tensor) -> torch.Tensor:
        feats = self.feature_extractor(audio, sampling_rate=16000, return_attention_mask=True, return_tensors="pt",
                                       padding="longest", pad_to_multiple_of=self.stride)
        return self.encoder(feats["input_features"].to(self.device).to(self.dtype),
                            feats["attention_mask"].to(self.device))
