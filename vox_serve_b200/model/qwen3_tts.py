"""Qwen3-TTS adapter on the sm_100a kernels: class name, properties and method contract of
``vox_serve/model/qwen3_tts.py:947-2045`` (``Qwen3TTSModel(BaseLMWithDepth)``) over ``depth_engine.Qwen3TTSEngine`` (talker +
code predictor, whole frame on the device) and ``tokenizer.qwen3_codec.Qwen3TTSDecoder`` (streaming 12 Hz codec decoder with a
per-request ``Qwen3TTSDecoderCache``).

What differs from the reference adapter:
  * ``frame_device`` runs a whole frame on the device (talker step -> codebook 0 -> the 15 predictor steps with their samplers
    -> the bf16 running sum of the predictor embeddings = the next row's ``input_features``); the reference's worker alternates
    16 graph replays with host glue per frame (cuda_graph_worker.py:1058-1160);
  * prompts arrive as pre-built rows: ``preprocess(prompt=(ids [T, 17], masks [T, 17], features [T, H]))`` -- column 0 the
    codebook-0 id, the last column the text id, ``masks[:, -1]`` = "this row also carries a codec embedding", features = the
    speaker embedding / ICL codec sums (qwen3_tts.py:1373-1803 builds them from text, a reference recording and the speaker
    encoder; tokenizer files and those checkpoints do not exist offline).  Text streaming (``is_input_streaming``) is not
    supported;
  * weights: ``state_dict=`` (+ ``audio_decoder_state_dict=``) or ``qwen3-tts-synthetic[-tiny][:seed]`` (seeded weights).
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, List, Optional

import torch

from .._lib import VoxB200Error
from ..depth_engine import CP, TK, Qwen3TTSDims, Qwen3TTSEngine, Qwen3TTSWeights
from ..engine import hf_layer_names
from ..requests import Request
from ..sampling import SamplingConfig
from .base import BaseLMWithDepth, PreprocessOutput

BF16 = torch.bfloat16

TINY = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
            intermediate_size=512, vocab_size=96, text_vocab_size=120, text_hidden_size=192, num_code_groups=6,
            cp_hidden_size=128, cp_num_hidden_layers=2, cp_num_attention_heads=2, cp_num_key_value_heads=1, cp_head_dim=64,
            cp_intermediate_size=256, cp_vocab_size=64, tts_pad_token_id=7)


def synthetic_state_dict(d: Qwen3TTSDims, seed: int = 0, device="cuda", head_scale: float = 8.0) -> Dict[str, torch.Tensor]:
    """Seeded bf16 weights under the reference's state_dict names (qwen3_tts.py:535-944), generated on ``device``."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)

    def rnd(*shape, std=0.02, mean=0.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float32, device=dev) * std + mean).to(BF16)

    sd: Dict[str, torch.Tensor] = {}

    def stack(prefix, n_layers, H, I, nq, nkv, D):
        for i in range(n_layers):
            n = hf_layer_names(i, prefix)
            sd[n["ln1"]], sd[n["ln2"]] = rnd(H, std=0.1, mean=1.0), rnd(H, std=0.1, mean=1.0)
            sd[n["q"]], sd[n["k"]], sd[n["v"]], sd[n["o"]] = rnd(nq * D, H), rnd(nkv * D, H), rnd(nkv * D, H), rnd(H, nq * D)
            sd[n["gate"]], sd[n["up"]], sd[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
            lp = f"{prefix}layers.{i}.self_attn."
            sd[lp + "q_norm.weight"], sd[lp + "k_norm.weight"] = rnd(D, std=0.1, mean=1.0), rnd(D, std=0.1, mean=1.0)
        sd[prefix + "norm.weight"] = rnd(H, std=0.1, mean=1.0)

    H, Hc, N = d.hidden_size, d.cp_hidden_size, d.num_code_groups
    sd[TK + "text_embedding.weight"] = rnd(d.text_vocab_size, d.text_hidden_size, std=1.0)
    sd[TK + "codec_embedding.weight"] = rnd(d.vocab_size, H, std=1.0)
    sd["talker.text_projection.linear_fc1.weight"] = rnd(d.text_hidden_size, d.text_hidden_size, std=d.text_hidden_size ** -0.5)
    sd["talker.text_projection.linear_fc1.bias"] = rnd(d.text_hidden_size, std=0.05)
    sd["talker.text_projection.linear_fc2.weight"] = rnd(H, d.text_hidden_size, std=d.text_hidden_size ** -0.5)
    sd["talker.text_projection.linear_fc2.bias"] = rnd(H, std=0.05)
    stack(TK, d.num_hidden_layers, H, d.intermediate_size, d.num_attention_heads, d.num_key_value_heads, d.head_dim)
    sd["talker.codec_head.weight"] = rnd(d.vocab_size, H, std=0.02 * head_scale)
    stack(CP, d.cp_num_hidden_layers, Hc, d.cp_intermediate_size, d.cp_num_attention_heads, d.cp_num_key_value_heads,
          d.cp_head_dim)
    sd["talker.code_predictor.small_to_mtp_projection.weight"] = rnd(Hc, H, std=0.05)
    sd["talker.code_predictor.small_to_mtp_projection.bias"] = rnd(Hc, std=0.05)
    for i in range(N - 1):
        sd[f"{CP}codec_embedding.{i}.weight"] = rnd(d.cp_vocab_size, H, std=1.0)
        sd[f"talker.code_predictor.lm_head.{i}.weight"] = rnd(d.cp_vocab_size, Hc, std=0.02 * head_scale)
    return sd


class Qwen3TTSModel(BaseLMWithDepth):
    def __init__(self, model_name, dtype=BF16, device="cuda:0", tokenizer_path=None, enable_torch_compile=False,
                 audio_decoder_device=None, state_dict: Optional[Dict] = None, dims: Optional[Qwen3TTSDims] = None,
                 max_tokens: Optional[int] = None, audio_decoder_state_dict: Optional[Dict] = None, codec_config=None,
                 detokenize_interval: Optional[int] = None, stop_token_id: Optional[int] = None,
                 suppress_tokens: Optional[List[int]] = None):
        if dtype != BF16:
            raise VoxB200Error("the B200 decode path computes in bf16 only")
        super().__init__(model_name, device, dtype, enable_torch_compile, audio_decoder_device)
        if not torch.cuda.is_available():
            raise VoxB200Error("Qwen3TTSModel needs a CUDA device: there is no CPU path")
        from ..tokenizer.qwen3_codec import Qwen3CodecConfig, Qwen3TTSDecoder

        synthetic = model_name.startswith("qwen3-tts-synthetic")
        seed = int(model_name.split(":")[1]) if ":" in model_name else 0
        tiny = model_name.split(":")[0].endswith("-tiny")
        if state_dict is None:
            if not synthetic:
                raise VoxB200Error(f"'{model_name}': no network in this build; pass state_dict=... or "
                                   "model_name='qwen3-tts-synthetic[-tiny][:seed]'")
            dims = dims or (Qwen3TTSDims(**TINY) if tiny else Qwen3TTSDims())
            state_dict = synthetic_state_dict(dims, seed, device)
        self.dims = dims or Qwen3TTSDims()
        self.weights = Qwen3TTSWeights(state_dict, self.dims, device)
        del state_dict
        if codec_config is None:
            codec_config = Qwen3CodecConfig(num_quantizers=self.dims.num_code_groups, codebook_size=self.dims.cp_vocab_size)
            if tiny:
                codec_config = dataclasses.replace(codec_config, latent_dim=64, codebook_dim=32, decoder_dim=128, hidden_size=32,
                                                   intermediate_size=64, head_dim=8, num_attention_heads=4, num_hidden_layers=2,
                                                   num_key_value_heads=2)
        if audio_decoder_state_dict is None:
            if not synthetic:
                raise VoxB200Error("pass audio_decoder_state_dict= (the Qwen3-TTS-Tokenizer-12Hz decoder weights)")
            audio_decoder_state_dict = _synthetic_codec_state_dict(codec_config, seed)
        self.audio_decoder = Qwen3TTSDecoder(device=audio_decoder_device or device, config=codec_config,
                                             state_dict=audio_decoder_state_dict)
        self.text_tokenizer = None
        # qwen3_tts.py:1081-1090: stop on the codec EOS id; the codec's special ids are never sampled for codebook 0
        self.stop_token_id = self.dims.vocab_size - 1 if stop_token_id is None else stop_token_id
        self.suppress_tokens = list(suppress_tokens or [])
        self._detokenize_interval = detokenize_interval if detokenize_interval is not None else 10
        self._max_tokens = max_tokens
        self.default_sampling_config = SamplingConfig(top_k=50, top_p=None, min_p=None, temperature=0.9,
                                                      repetition_penalty=None, repetition_window=None, cfg_scale=None)
        self._engines: Dict[Any, Qwen3TTSEngine] = {}
        self._suppress_idx = None
        self.max_batch, self.max_rows = 64, 1024 + 64

    # ---- static facts (qwen3_tts.py:1096-1262) -------------------------------------------------------------
    needs_input_masks = True
    needs_input_features = True
    decode_text_column_mask = True       # decode rows: masks all ones (qwen3_tts.py:1941)
    detokenize_overlap = 0
    n_channels = 1

    @property
    def detokenize_interval(self) -> int:
        return self._detokenize_interval

    @property
    def output_audio_length(self) -> int:
        return self._detokenize_interval * self.audio_decoder.cfg.hop

    @property
    def n_codebooks(self) -> int:
        return self.dims.num_code_groups + 1

    @property
    def depth_n_codebooks(self) -> int:
        return self.dims.num_code_groups

    num_attention_heads = property(lambda self: self.dims.num_attention_heads)
    num_key_value_heads = property(lambda self: self.dims.num_key_value_heads)
    num_hidden_layers = property(lambda self: self.dims.num_hidden_layers)
    hidden_size = property(lambda self: self.dims.hidden_size)
    head_dim = property(lambda self: self.dims.head_dim)
    depth_num_attention_heads = property(lambda self: self.dims.cp_num_attention_heads)
    depth_num_key_value_heads = property(lambda self: self.dims.cp_num_key_value_heads)
    depth_num_hidden_layers = property(lambda self: self.dims.cp_num_hidden_layers)
    depth_hidden_size = property(lambda self: self.dims.cp_hidden_size)
    depth_head_dim = property(lambda self: self.dims.cp_head_dim)
    vocab_size = property(lambda self: self.dims.vocab_size)
    depth_vocab_size = property(lambda self: self.dims.cp_vocab_size)

    @property
    def max_tokens(self) -> int:
        if self.default_sampling_config.max_tokens is not None:
            return self.default_sampling_config.max_tokens
        return self._max_tokens if self._max_tokens is not None else 2048

    def is_stop_id(self, token_ids) -> bool:
        t = token_ids[0] if isinstance(token_ids, (list, tuple)) else token_ids
        return int(t) == self.stop_token_id

    def audio_decoder_initial_cache(self, batch_size: int):
        """qwen3_tts.py:1244-1259"""
        return self.audio_decoder.init_cache(batch_size, detokenize_interval=self.detokenize_interval)

    # ---- prompt side --------------------------------------------------------------------------------------
    def preprocess(self, prompt=None, audio_path: str = None, **kwargs) -> PreprocessOutput:
        assert audio_path is None
        if kwargs.get("is_input_streaming"):
            raise VoxB200Error("text streaming into a running request is not supported")
        if not (isinstance(prompt, (tuple, list)) and len(prompt) == 3):
            raise VoxB200Error("no tokenizer / speaker encoder offline: pass prompt=(ids [T, n_codebooks], masks [T, n_codebooks], "
                               "features [T, hidden])")
        ids = torch.as_tensor(prompt[0], dtype=torch.int64)
        masks = torch.as_tensor(prompt[1], dtype=torch.bool)
        feats = torch.as_tensor(prompt[2]).to(BF16)
        assert ids.shape == masks.shape and ids.shape[1] == self.n_codebooks and feats.shape == (ids.shape[0], self.hidden_size)
        return PreprocessOutput(input_tokens=ids, input_masks=masks, input_features=feats, repetition_cache=None)

    # ---- engine -------------------------------------------------------------------------------------------
    def engine_for(self, kv_cache: torch.Tensor, page_size: Optional[int] = None) -> Qwen3TTSEngine:
        key = (kv_cache.data_ptr(), tuple(kv_cache.shape))
        e = self._engines.get(key)
        if e is None:
            e = self._engines[key] = Qwen3TTSEngine(self.weights, kv_cache, page_size or kv_cache.shape[3],
                                                    max_batch=self.max_batch, max_rows=self.max_rows)
            if self.suppress_tokens:
                self._install_suppression(e)
        return e

    def _install_suppression(self, eng: Qwen3TTSEngine) -> None:
        """``logits[..., suppress_tokens] = finfo.min`` before codebook 0 is sampled (qwen3_tts.py:1894-1895)."""
        idx = torch.tensor(self.suppress_tokens, dtype=torch.int64, device=eng.device)
        tail = eng.frame_tail

        def frame_tail(B, logits0, hidden, cfg, keep_logits=None):
            logits0.index_fill_(1, idx, torch.finfo(logits0.dtype).min)
            return tail(B, logits0, hidden, cfg, keep_logits)

        eng.frame_tail = frame_tail

    def frame_device(self, kv_cache: torch.Tensor, attn_wrapper, position_ids: torch.Tensor, n_req: int,
                     input_ids: Optional[torch.Tensor] = None, input_masks: Optional[torch.Tensor] = None,
                     last_rows: Optional[torch.Tensor] = None, sampling_params: Optional[SamplingConfig] = None,
                     input_features: Optional[torch.Tensor] = None):
        """One whole frame for ``n_req`` requests on the device -> ids [n_req, n_codebooks] int64 (the last column is the
        text stream: tts_pad, qwen3_tts.py:1916).  Decode: ``input_ids`` [n_req, n_codebooks] (only codebook 0 is read) and
        ``input_features`` [n_req, H] = the predictor-embedding sums left by the previous frames.  Prefill: the concatenated
        prompt rows with ``input_masks`` / ``input_features`` and ``last_rows``.  The new sums are in ``engine.feat[:n_req]``."""
        eng = self.engine_for(kv_cache, attn_wrapper.page_size)
        cfg = sampling_params or self.default_sampling_config
        N = self.dims.num_code_groups
        if last_rows is not None:
            out = eng.prefill_frame(input_ids[:, -1], input_ids[:, 0], input_masks[:, -1], input_features, position_ids,
                                    last_rows, attn_wrapper.plan_rows, cfg)
        else:
            eng.frame[0, :n_req].copy_(input_ids[:, 0])
            eng.feat[:n_req].copy_(input_features)
            out = eng.decode_frame(n_req, position_ids, attn_wrapper.plan_rows, cfg)
        pad = torch.full((n_req, 1), self.dims.tts_pad_token_id, dtype=torch.int64, device=out.device)
        return torch.cat([out[:, :N], pad], dim=1)

    def slot_features(self, kv_cache: torch.Tensor, n_req: int) -> torch.Tensor:
        """input_features of the NEXT row of the n_req requests of the last ``frame_device`` call."""
        return self.engine_for(kv_cache).feat[:n_req]

    # the step-at-a-time surface of the reference (forward / sampling / depth_forward / depth_sampling) is what frame_device
    # replaces; the CSM adapter keeps both, here only the fused form exists
    def forward(self, *a, **k):
        raise VoxB200Error("Qwen3TTSModel runs whole frames: use frame_device (the worker does)")

    sampling = depth_forward = depth_sampling = forward

    # ---- vocoder ------------------------------------------------------------------------------------------
    def postprocess(self, token_ids: torch.Tensor, decoder_cache=None, **kwargs) -> torch.Tensor:
        """[B, interval, n_codebooks] frame rows -> [B, 1, interval * 1920] (qwen3_tts.py:2006-2044): the text column is
        dropped; with a cache the chunk continues the streams' codec state (updated in place), without one it is decoded
        from a zero state."""
        codes = token_ids[:, :, :-1].transpose(1, 2).clamp(0, self.audio_decoder.cfg.codebook_size - 1)
        wav, _ = self.audio_decoder.decode_chunk(codes, decoder_cache)
        return wav


def _synthetic_codec_state_dict(cfg, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded decoder weights under the checkpoint's key names (scales keep every stage O(1))."""
    import math

    g = torch.Generator().manual_seed(seed + 101)

    def rnd(*shape, std=1.0, mean=0.0):
        return torch.randn(*shape, generator=g) * std + mean

    sd: Dict[str, torch.Tensor] = {}
    dq, C, Lt, H = cfg.codebook_dim // 2, cfg.codebook_dim, cfg.latent_dim, cfg.hidden_size
    for name, n in (("rvq_first", 1), ("rvq_rest", cfg.num_quantizers - 1)):
        for k in range(n):
            p = f"quantizer.{name}.vq.layers.{k}._codebook."
            sd[p + "cluster_usage"] = torch.rand(cfg.codebook_size, generator=g) + 0.5
            sd[p + "embedding_sum"] = rnd(cfg.codebook_size, dq) * sd[p + "cluster_usage"][:, None]
        sd[f"quantizer.{name}.output_proj.weight"] = rnd(C, dq, 1, std=1.0 / math.sqrt(dq * cfg.num_quantizers))
    sd["pre_conv.conv.weight"], sd["pre_conv.conv.bias"] = rnd(Lt, C, 3, std=(3 * C) ** -0.5), rnd(Lt, std=0.05)
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
