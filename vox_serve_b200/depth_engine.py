"""Backbone + depth-transformer frame step on the sm_100a kernels.

The reference runs one audio frame of a multi-codebook model as 1 backbone graph replay + (N - 1) depth graph replays,
each with a FlashInfer plan, two device synchronisations and a ``.item()`` per request between them
(``vox_serve/worker/cuda_graph_worker.py:1058-1160``, ``vox_serve/model/csm.py:665-769``).  Here the whole frame --
frame embedding, backbone step, codebook-0 sample, the 2-row depth prefill and the N - 2 one-row depth decodes with
their samples and embedding look-ups -- is ONE launch sequence over device-resident indices: CUDA-graph capturable, no
host round trip inside a frame.

* the depth decoder's paged-KV page table is the same every frame (request j owns page j of a per-frame cache of
  ``n_codebooks`` slots, cuda_graph_worker.py:1070-1076), so its row plans are computed once per batch size;
* per-position output heads (``CsmCodebooksHead``: row at depth position p uses ``weight[p - 1]``, csm.py:235-255) are
  packed once as N - 1 projection weights; every row of a depth step shares its position, so each step is one GEMM;
* the frame lives codebook-major on the device (``frame[c][b]``): every sampler call writes one contiguous row, and the
  next embedding look-up reads it back without leaving the device.

The depth cache is NOT zeroed per frame (the reference does, :1076): attention only reads the ``kv_len`` slots the
frame has already written.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._lib import VoxB200Error
from .engine import LlamaDims, LlamaEngine, LlamaWeights
from .sampling import SamplingConfig, strategy_of

BF16 = torch.bfloat16


@dataclass
class CsmDims:
    """transformers.CsmConfig / CsmDepthDecoderConfig fields the path needs (defaults = sesame/csm-1b)."""
    hidden_size: int = 2048
    num_hidden_layers: int = 16
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 64
    intermediate_size: int = 8192
    num_codebooks: int = 32
    vocab_size: int = 2051
    text_vocab_size: int = 128256
    depth_hidden_size: int = 1024
    depth_num_hidden_layers: int = 4
    depth_num_attention_heads: int = 8
    depth_num_key_value_heads: int = 2
    depth_head_dim: int = 128
    depth_intermediate_size: int = 8192
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0

    def backbone(self) -> LlamaDims:      # csm.py:66-70: llama-3.1 smoothing defaults factor 32 / 1 / 4 / 8192
        return LlamaDims(self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.num_key_value_heads,
                         self.head_dim, self.intermediate_size, self.vocab_size, self.rms_norm_eps, self.rope_theta)

    def depth(self) -> LlamaDims:
        return LlamaDims(self.depth_hidden_size, self.depth_num_hidden_layers, self.depth_num_attention_heads,
                         self.depth_num_key_value_heads, self.depth_head_dim, self.depth_intermediate_size,
                         self.vocab_size, self.rms_norm_eps, self.rope_theta)


BB, DD = "backbone_model.", "depth_decoder.model."


class CsmWeights:
    """Device weights of ``CsmForConditionalGeneration`` under the reference's state_dict names (csm.py:158-312)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: CsmDims, device="cuda"):
        dev = torch.device(device)
        self.dims = dims

        def put(t):
            return t.to(device=dev, dtype=BF16).contiguous()

        self.backbone = LlamaWeights.from_state_dict(sd, dims.backbone(), dev, prefix=BB, embed_key=None,
                                                     head_key="lm_head.weight")
        head = sd["depth_decoder.codebooks_head.weight"]                    # [N - 1, H_depth, vocab]
        assert tuple(head.shape) == (dims.num_codebooks - 1, dims.depth_hidden_size, dims.vocab_size), tuple(head.shape)
        self.depth = LlamaWeights.from_state_dict(sd, dims.depth(), dev, prefix=DD, embed_key=None,
                                                  heads=[head[p].t().contiguous() for p in range(head.shape[0])])
        self.embed_audio = put(sd[BB + "embed_tokens.embed_audio_tokens.weight"])       # [N * vocab, H]
        self.embed_text = put(sd["embed_text_tokens.weight"])                            # [text_vocab, H]
        self.projector = ops.pack_weight(put(sd[DD + "inputs_embeds_projector.weight"]), 128)   # [H_depth, H]
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
            torch.cuda.empty_cache()


class DepthFrameEngine:
    """Shared frame machinery of the backbone + depth-decoder models: static depth plans, the codebook-major frame
    buffer, the sampler glue and the depth loop.  Subclasses say where the depth decoder's inputs come from."""

    def __init__(self, backbone_w: LlamaWeights, depth_w: LlamaWeights, projector, projector_bias, n_codebooks: int,
                 kv_cache: torch.Tensor, page_size: int, max_batch: int, max_rows: int, frame_rows: int):
        dev = kv_cache.device
        self.device, self.max_batch, self.n_cb = dev, max_batch, n_codebooks
        dd = depth_w.dims
        self.bb = LlamaEngine(backbone_w, kv_cache, page_size, max_rows=max_rows)
        self.depth_page = ((n_codebooks + 15) // 16) * 16                     # attention tiles are multiples of 16 tokens
        self.depth_kv = torch.zeros(dd.num_hidden_layers, max_batch, 2, self.depth_page, dd.num_key_value_heads, dd.head_dim,
                                    dtype=BF16, device=dev)
        self.dp = LlamaEngine(depth_w, self.depth_kv, self.depth_page, max_rows=2 * max_batch)
        self.projector, self.projector_bias, self.depth_heads = projector, projector_bias, depth_w.heads
        H = backbone_w.dims.hidden_size
        # frame[c][b]: ids of the frame being generated / fed back
        self.frame = torch.zeros(frame_rows, max_batch, dtype=torch.int64, device=dev)
        self.out_ids = torch.zeros(max_batch, frame_rows, dtype=torch.int64, device=dev)
        self.embed_tmp = torch.zeros(max_batch, H, dtype=BF16, device=dev)
        self.pair_tmp = torch.zeros(2 * max_batch, H, dtype=BF16, device=dev)
        self._plans: Dict[int, List[Tuple[ops.RowPlan, torch.Tensor, Optional[torch.Tensor]]]] = {}
        seed = torch.cuda.default_generators[dev.index or 0].initial_seed() & ((1 << 62) - 1)
        self.rng_state = torch.tensor([seed, 0, 0], dtype=torch.int64, device=dev)

    @property
    def max_rows(self) -> int:
        return self.bb.max_rows

    # ---- static depth-decoder plans ---------------------------------------------------------------------
    def depth_plans(self, B: int):
        """[(row plan, positions, last_rows)] for depth steps 1 .. N-1 at batch B (cuda_graph_worker.py:1070-1160:
        step 1 = 2-row prefill at positions 0, 1; step i >= 2 = 1-row decode at position i, kv length i + 1)."""
        if B in self._plans:
            return self._plans[B]
        dev, N = self.device, self.n_cb
        i32 = dict(dtype=torch.int32, device=dev)
        kv_indptr, kv_indices = torch.arange(B + 1, **i32), torch.arange(B, **i32)
        plans = []
        for i in range(1, N):
            plan = ops.RowPlan(2 * B if i == 1 else B, dev)
            if i == 1:
                qo = torch.arange(B + 1, **i32) * 2
                ops.plan_rows(plan, qo, kv_indptr, kv_indices, torch.full((B,), 2, **i32), B, 2 * B, self.depth_page,
                              self.dp.chunk)
                pos = torch.tensor([0, 1] * B, **i32)
                last = (torch.arange(B, **i32) * 2 + 1).contiguous()
            else:
                ops.plan_rows(plan, None, kv_indptr, kv_indices, torch.full((B,), i + 1, **i32), B, B, self.depth_page,
                              self.dp.chunk)
                pos, last = torch.full((B,), i, **i32), None
            plans.append((plan, pos, last))
        torch.cuda.synchronize(dev)
        self._plans[B] = plans
        return plans

    def _sample(self, logits: torch.Tensor, cfg: SamplingConfig, out: torch.Tensor) -> None:
        kind = strategy_of(cfg)
        ops.sample(logits, kind, top_k=cfg.top_k or 0, top_p=1.0 if cfg.top_p is None else cfg.top_p,
                   min_p=cfg.min_p or 0.0, temperature=cfg.temperature if kind != "greedy" else 1.0,
                   rng_state=self.rng_state, out=out)

    # ---- hooks ------------------------------------------------------------------------------------------
    def _embed_cb0(self, B: int, out: torch.Tensor) -> None:
        """embedding of the just-sampled codebook 0 (second row of the depth prefill) -> out [B, H]"""
        raise NotImplementedError

    def _embed_step(self, i: int, B: int, out: torch.Tensor) -> None:
        """input of depth step i >= 2: embedding of codebook i - 1 -> out [B, H]"""
        raise NotImplementedError

    def _finish_frame(self, B: int) -> None:
        pass

    # ---- one frame ----------------------------------------------------------------------------------------
    def frame_tail(self, B: int, logits0: torch.Tensor, hidden: torch.Tensor, cfg: SamplingConfig,
                   keep_logits: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """Codebook 0 from the backbone logits [B, vocab], codebooks 1 .. N-1 from the depth decoder started from the
        backbone state ``hidden`` [B, H] (csm.py:665-769 / qwen3_tts.py:1863-2004 + cuda_graph_worker.py:1058-1160).
        Leaves the frame in ``self.frame[:, :B]`` and returns the request-major ids [B, frame rows].  ``keep_logits``:
        a list that receives a copy of every step's logits (tests)."""
        N, fr = self.n_cb, self.frame
        self._sample(logits0, cfg, fr[0, :B])
        if keep_logits is not None:
            keep_logits.append(logits0.clone())
        e = self.embed_tmp[:B]
        self._embed_cb0(B, e)
        pair = ops.interleave_rows(hidden, e, out=self.pair_tmp[:2 * B])
        plans = self.depth_plans(B)
        for i in range(1, N):
            plan, pos, last = plans[i - 1]
            R = 2 * B if i == 1 else B
            if i > 1:
                self._embed_step(i, B, e)
            ops.gemm(pair if i == 1 else e, self.projector, mode=0, out=self.dp.hidden[:R], bias=self.projector_bias)
            logits = self.dp.forward(None, pos, R, last_rows=last, plan=plan, head=self.depth_heads[i - 1])
            self._sample(logits[:B], cfg, fr[i, :B])
            if keep_logits is not None:
                keep_logits.append(logits[:B].clone())
        self._finish_frame(B)
        return ops.transpose_i64(fr[:, :B], out=self.out_ids[:B])


class CsmEngine(DepthFrameEngine):
    """One replica's CSM backbone + depth decoder over a backbone KV cache ``[L, pages, 2, page, Hkv, D]``."""

    def __init__(self, weights: CsmWeights, kv_cache: torch.Tensor, page_size: int, max_batch: int, max_rows: int):
        d = weights.dims
        self.w, self.dims = weights, d
        super().__init__(weights.backbone, weights.depth, weights.projector, None, d.num_codebooks, kv_cache, page_size,
                         max_batch, max_rows, frame_rows=d.num_codebooks + 1)          # last row = the text stream

    # ---- frame inputs -----------------------------------------------------------------------------------
    def embed_prompt(self, ids: torch.Tensor, masks: torch.Tensor, row0: int = 0) -> int:
        """Prompt rows ``ids`` / ``masks`` [T, N + 1] (int64 / bool, device; last column = text stream) -> backbone
        input rows ``bb.hidden[row0 : row0 + T]`` (csm.py:647-654)."""
        T, N = ids.shape[0], self.dims.num_codebooks
        ops.multi_embed_sum(self.bb.hidden[row0:row0 + T], ids, self.w.embed_audio, col_offset=self.dims.vocab_size,
                            n_cols_a=N, table_b=self.w.embed_text, mask=masks)
        return T

    def embed_frames(self, B: int, row0: int = 0) -> None:
        """The previous frame of every request (``self.frame[:N, :B]``, text stream masked off as in csm.py:711-712)
        -> backbone input rows ``bb.hidden[row0 : row0 + B]``."""
        N = self.dims.num_codebooks
        ops.multi_embed_sum(self.bb.hidden[row0:row0 + B], self.frame[:N, :B].t(), self.w.embed_audio,
                            col_offset=self.dims.vocab_size)

    def _embed_cb0(self, B, out):
        ops.multi_embed_sum(out, self.frame[0:1, :B].t(), self.w.embed_audio, col_offset=self.dims.vocab_size, col0=0)

    def _embed_step(self, i, B, out):       # codebook i - 1 through the BACKBONE's audio table (csm.py:760-761)
        ops.multi_embed_sum(out, self.frame[i - 1:i, :B].t(), self.w.embed_audio, col_offset=self.dims.vocab_size, col0=i - 1)

    def _finish_frame(self, B):             # text row = codebook 0: the reference's ``repeat`` quirk (csm.py:693)
        N = self.dims.num_codebooks
        self.frame[N, :B].copy_(self.frame[0, :B])

    def decode_frame(self, B: int, position_ids: torch.Tensor, plan: ops.RowPlan, cfg: SamplingConfig,
                     keep_logits=None) -> torch.Tensor:
        """One decode step for B running requests: feed back ``self.frame``, run the backbone at ``position_ids``
        (int32 [B]) under the backbone row plan, then the frame tail.  No host interaction: capturable."""
        self.embed_frames(B)
        logits0, hidden = self.bb.forward(None, position_ids, B, plan=plan, want_hidden=True)
        return self.frame_tail(B, logits0, hidden, cfg, keep_logits)

    def prefill_frame(self, ids: torch.Tensor, masks: torch.Tensor, position_ids: torch.Tensor, last_rows: torch.Tensor,
                      plan: ops.RowPlan, cfg: SamplingConfig, keep_logits=None) -> torch.Tensor:
        """Prompt rows of B requests (ragged, concatenated) -> their first frame.  ``last_rows`` int32 [B]: index of
        every request's last prompt row (qo_indptr[1:] - 1)."""
        T, B = ids.shape[0], last_rows.numel()
        if B > self.max_batch:
            raise VoxB200Error(f"{B} requests exceed the engine's max_batch {self.max_batch}")
        self.embed_prompt(ids, masks)
        logits0, hidden = self.bb.forward(None, position_ids, T, last_rows=last_rows, plan=plan, want_hidden=True)
        return self.frame_tail(B, logits0, hidden, cfg, keep_logits)


# =====================================================================================================================
# Qwen3-TTS: talker + code predictor (vox_serve/model/qwen3_tts.py:535-944, 1805-2004)
# =====================================================================================================================
@dataclass
class Qwen3TTSDims:
    """Qwen3TTSTalkerConfig / Qwen3TTSCodePredictorConfig fields the path needs (qwen3_tts.py:112-253 defaults)."""
    hidden_size: int = 2048
    num_hidden_layers: int = 28
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 6144
    vocab_size: int = 3072
    text_vocab_size: int = 151936
    text_hidden_size: int = 2048
    num_code_groups: int = 16
    cp_hidden_size: int = 1024
    cp_num_hidden_layers: int = 5
    cp_num_attention_heads: int = 16
    cp_num_key_value_heads: int = 8
    cp_head_dim: int = 128
    cp_intermediate_size: int = 3072
    cp_vocab_size: int = 2048
    rms_norm_eps: float = 1e-6
    rope_theta: float = 1000000.0
    tts_pad_token_id: int = 151671

    def talker(self) -> LlamaDims:       # plain RoPE, q/k RMSNorm per head (qwen3_tts.py:578-653)
        return LlamaDims(self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.num_key_value_heads,
                         self.head_dim, self.intermediate_size, self.vocab_size, self.rms_norm_eps, self.rope_theta, 1.0,
                         None, None, None, qk_norm=True)

    def predictor(self) -> LlamaDims:
        return LlamaDims(self.cp_hidden_size, self.cp_num_hidden_layers, self.cp_num_attention_heads,
                         self.cp_num_key_value_heads, self.cp_head_dim, self.cp_intermediate_size, self.cp_vocab_size,
                         self.rms_norm_eps, self.rope_theta, 1.0, None, None, None, qk_norm=True)


TK, CP = "talker.model.", "talker.code_predictor.model."


class Qwen3TTSWeights:
    """Device weights of the talker + code predictor under the reference's state_dict names (qwen3_tts.py:707-833)."""

    def __init__(self, sd: Dict[str, torch.Tensor], dims: Qwen3TTSDims, device="cuda"):
        dev = torch.device(device)
        self.dims = dims

        def put(t):
            return t.to(device=dev, dtype=BF16).contiguous()

        N = dims.num_code_groups
        self.talker = LlamaWeights.from_state_dict(sd, dims.talker(), dev, prefix=TK, embed_key=None,
                                                   head_key="talker.codec_head.weight")
        self.predictor = LlamaWeights.from_state_dict(
            sd, dims.predictor(), dev, prefix=CP, embed_key=None,
            heads=[sd[f"talker.code_predictor.lm_head.{i}.weight"] for i in range(N - 1)])
        self.codec_embedding = put(sd[TK + "codec_embedding.weight"])                   # [vocab, H]
        # the predictor's per-codebook tables stacked: row (i - 1) * cp_vocab + id embeds codebook i
        self.cp_embedding = put(torch.cat([sd[f"{CP}codec_embedding.{i}.weight"] for i in range(N - 1)], 0))
        self.projector = ops.pack_weight(put(sd["talker.code_predictor.small_to_mtp_projection.weight"]), 128)
        self.projector_bias = put(sd["talker.code_predictor.small_to_mtp_projection.bias"])
        # prompt-side text path (qwen3_tts.py:656-664): text embedding -> 2-layer SiLU MLP with bias
        self.text_embedding = put(sd[TK + "text_embedding.weight"])
        self.fc1_w, self.fc1_b = put(sd["talker.text_projection.linear_fc1.weight"]), put(sd["talker.text_projection.linear_fc1.bias"])
        self.fc2_w, self.fc2_b = put(sd["talker.text_projection.linear_fc2.weight"]), put(sd["talker.text_projection.linear_fc2.bias"])
        # every decode row carries the projected embedding of tts_pad (qwen3_tts.py:1934-1946): a constant
        self.text_pad = self.text_project(torch.tensor([dims.tts_pad_token_id], device=dev))[0].contiguous()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
            torch.cuda.empty_cache()

    def text_project(self, text_ids: torch.Tensor) -> torch.Tensor:
        """text ids [T] -> [T, H] bf16.  Prompt side only (a few dozen rows once per request): library GEMMs."""
        import torch.nn.functional as F

        t = F.embedding(text_ids.long(), self.text_embedding)
        return F.linear(F.silu(F.linear(t, self.fc1_w, self.fc1_b)), self.fc2_w, self.fc2_b)


class Qwen3TTSEngine(DepthFrameEngine):
    """Talker + code predictor over a talker KV cache.  frame[0] = codebook 0 (talker vocabulary), frame[1 .. N-1] = the
    predictor's codebooks; ``self.feat`` [B, H] = the bf16 running sum of the predictor embeddings of the last frame,
    the next talker row's ``input_features`` (qwen3_tts.py:1981-2004)."""

    def __init__(self, weights: Qwen3TTSWeights, kv_cache: torch.Tensor, page_size: int, max_batch: int, max_rows: int):
        d = weights.dims
        self.w, self.dims = weights, d
        super().__init__(weights.talker, weights.predictor, weights.projector, weights.projector_bias, d.num_code_groups,
                         kv_cache, page_size, max_batch, max_rows, frame_rows=d.num_code_groups)
        self.feat = torch.zeros(max_batch, d.hidden_size, dtype=BF16, device=kv_cache.device)

    def _embed_cb0(self, B, out):            # the TALKER's codec embedding of codebook 0 (qwen3_tts.py:1925-1927)
        ops.multi_embed_sum(out, self.frame[0:1, :B].t(), self.w.codec_embedding)

    def _embed_step(self, i, B, out):        # code_predictor.codec_embedding[i - 2] of codebook i - 1 (qwen3_tts.py:1996-2000)
        ops.multi_embed_sum(out, self.frame[i - 1:i, :B].t(), self.w.cp_embedding, col_offset=self.dims.cp_vocab_size,
                            col0=i - 2)

    def _finish_frame(self, B):              # input_features of the next talker row: e_1 + e_2 + ... in bf16, in order
        N = self.dims.num_code_groups
        ops.multi_embed_sum(self.feat[:B], self.frame[1:N, :B].t(), self.w.cp_embedding, col_offset=self.dims.cp_vocab_size,
                            round_each=True)

    def decode_frame(self, B: int, position_ids: torch.Tensor, plan: ops.RowPlan, cfg: SamplingConfig,
                     keep_logits=None) -> torch.Tensor:
        """Next frame of B running requests from ``self.frame[0]`` / ``self.feat`` left by the previous call."""
        ops.talker_embed(self.bb.hidden[:B], self.w.text_pad, self.w.codec_embedding, self.frame[0, :B], None, self.feat[:B])
        logits0, hidden = self.bb.forward(None, position_ids, B, plan=plan, want_hidden=True)
        return self.frame_tail(B, logits0, hidden, cfg, keep_logits)

    def prefill_frame(self, text_ids: torch.Tensor, cb0: torch.Tensor, needs_codec: torch.Tensor, features: torch.Tensor,
                      position_ids: torch.Tensor, last_rows: torch.Tensor, plan: ops.RowPlan, cfg: SamplingConfig,
                      keep_logits=None) -> torch.Tensor:
        """Prompt rows (text id, codebook-0 id, needs_codec, input_features [T, H]) of B requests -> their first frame."""
        T, B = text_ids.shape[0], last_rows.numel()
        if B > self.max_batch:
            raise VoxB200Error(f"{B} requests exceed the engine's max_batch {self.max_batch}")
        text = self.w.text_project(text_ids)
        ops.talker_embed(self.bb.hidden[:T], text, self.w.codec_embedding, cb0.to(torch.int64).contiguous(),
                         needs_codec.contiguous(), features.contiguous())
        logits0, hidden = self.bb.forward(None, position_ids, T, last_rows=last_rows, plan=plan, want_hidden=True)
        return self.frame_tail(B, logits0, hidden, cfg, keep_logits)
