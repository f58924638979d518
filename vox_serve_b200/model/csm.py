"""CSM-1B adapter on the sm_100a kernels: the class name, properties and method contract of
``vox_serve/model/csm.py:315-789`` (``CSMModel(BaseLMWithDepth)``) over ``depth_engine.CsmEngine``.

What differs from the reference adapter:
  * ``frame_device`` runs a whole frame on the device (backbone step -> codebook 0 -> 31 depth steps), where the
    reference's worker alternates 32 graph replays with host-side sampling glue (cuda_graph_worker.py:1058-1160);
    ``forward`` / ``sampling`` / ``depth_forward`` / ``depth_sampling`` remain available one step at a time with the
    reference's signatures for callers that drive the loop themselves;
  * prompts arrive as pre-built frame rows: ``preprocess(prompt=(ids [T, 33], masks [T, 33]))``.  The reference tokenises
    text with the Llama-3.2 tokenizer and prepends two built-in voice prompts encoded with the Mimi ENCODER
    (csm.py:511-568, 613-615); neither the tokenizer files nor that checkpoint exist offline;
  * weights: a local HF directory or ``state_dict=``, or ``csm-synthetic[-tiny][:seed]`` (seeded weights);
  * ``postprocess`` runs ``tokenizer.mimi.MimiDecoder`` (csrc/mimi.cu); its weights come from a local safetensors file,
    ``audio_decoder_state_dict=`` or, for ``csm-synthetic`` names, a seeded generator.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, List, Optional, Tuple

import torch

from .. import ops
from .._lib import VoxB200Error
from ..depth_engine import BB, DD, CsmDims, CsmEngine, CsmWeights
from ..engine import hf_layer_names
from ..requests import Request
from ..sampling import Sampler, SamplingConfig
from .base import BaseLMWithDepth, PreprocessOutput

BF16 = torch.bfloat16


def synthetic_state_dict(dims: CsmDims, seed: int = 0, device="cuda", head_scale: float = 8.0) -> Dict[str, torch.Tensor]:
    """Seeded bf16 weights under the reference's state_dict names (csm.py:158-312), generated on ``device``."""
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
        sd[prefix + "norm.weight"] = rnd(H, std=0.1, mean=1.0)

    H, Hd, N, V = dims.hidden_size, dims.depth_hidden_size, dims.num_codebooks, dims.vocab_size
    sd[BB + "embed_tokens.embed_audio_tokens.weight"] = rnd(N * V, H, std=1.0)
    stack(BB, dims.num_hidden_layers, H, dims.intermediate_size, dims.num_attention_heads, dims.num_key_value_heads,
          dims.head_dim)
    sd["lm_head.weight"] = rnd(V, H, std=0.02 * head_scale)
    sd["embed_text_tokens.weight"] = rnd(dims.text_vocab_size, H, std=1.0)
    stack(DD, dims.depth_num_hidden_layers, Hd, dims.depth_intermediate_size, dims.depth_num_attention_heads,
          dims.depth_num_key_value_heads, dims.depth_head_dim)
    sd[DD + "inputs_embeds_projector.weight"] = rnd(Hd, H, std=0.05)
    sd["depth_decoder.codebooks_head.weight"] = rnd(N - 1, Hd, V, std=0.02 * head_scale)
    return sd


TINY = dict(hidden_size=512, num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=2, head_dim=64,
            intermediate_size=1024, num_codebooks=8, vocab_size=515, text_vocab_size=1000, depth_hidden_size=256,
            depth_num_hidden_layers=2, depth_num_attention_heads=2, depth_num_key_value_heads=1, depth_head_dim=128,
            depth_intermediate_size=512)


class CSMModel(BaseLMWithDepth):
    def __init__(self, model_name, dtype=BF16, device="cuda:0", tokenizer_path="meta-llama/Llama-3.2-1B",
                 enable_torch_compile=False, audio_decoder_device=None, state_dict: Optional[Dict] = None,
                 dims: Optional[CsmDims] = None, max_tokens: Optional[int] = None,
                 audio_decoder_state_dict: Optional[Dict] = None, mimi_config=None):
        if model_name == "csm":
            model_name = "sesame/csm-1b"
        if dtype != BF16:
            raise VoxB200Error("the B200 decode path computes in bf16 only")
        super().__init__(model_name, device, dtype, enable_torch_compile, audio_decoder_device)
        if not torch.cuda.is_available():
            raise VoxB200Error("CSMModel needs a CUDA device: there is no CPU path")
        if state_dict is None:
            if model_name.startswith("csm-synthetic"):
                seed = int(model_name.split(":")[1]) if ":" in model_name else 0
                tiny = model_name.split(":")[0].endswith("-tiny")
                dims = dims or (CsmDims(**TINY) if tiny else CsmDims())
                state_dict = synthetic_state_dict(dims, seed, device)
            else:
                raise VoxB200Error(f"'{model_name}': no network in this build; pass state_dict=... or "
                                   "model_name='csm-synthetic[-tiny][:seed]'")
        self.dims = dims or CsmDims()
        self.weights = CsmWeights(state_dict, self.dims, device)
        del state_dict
        self.text_tokenizer = None
        from ..tokenizer.mimi import MimiConfig, MimiDecoder, synthetic_state_dict as mimi_synthetic

        synthetic = model_name.startswith("csm-synthetic")
        if mimi_config is None:
            mimi_config = (MimiConfig(dimension=64, n_filters=8, codebook_dim=32, num_heads=2, num_layers=2, dim_feedforward=128)
                           if synthetic and model_name.split(":")[0].endswith("-tiny") else MimiConfig())
        if audio_decoder_state_dict is None and synthetic:
            audio_decoder_state_dict = mimi_synthetic(mimi_config, int(model_name.split(":")[1]) if ":" in model_name else 0)
        # csm.py:349-353: MimiDecoder(model_repo="kyutai/moshiko-pytorch-bf16", ..., num_codebooks=32)
        self.audio_decoder = MimiDecoder(num_codebooks=self.dims.num_codebooks, mimi_config=mimi_config,
                                         device=audio_decoder_device or device, state_dict=audio_decoder_state_dict)
        self.stop_token_id = 0               # csm.py:355
        self._max_tokens = max_tokens
        self.default_sampling_config = SamplingConfig(top_k=50, top_p=None, min_p=None, temperature=0.9,
                                                      repetition_penalty=None, repetition_window=None, cfg_scale=None)
        self._engines: Dict[Any, CsmEngine] = {}
        self.max_batch, self.max_rows = 64, 1024 + 64

    # ---- static facts (csm.py:367-445, 571-612) ----------------------------------------------------------
    needs_input_masks = True
    detokenize_interval = 10
    detokenize_overlap = 0
    n_channels = 1
    output_audio_length = 19200

    @property
    def n_codebooks(self) -> int:
        return self.dims.num_codebooks + 1

    @property
    def depth_n_codebooks(self) -> int:
        return self.dims.num_codebooks

    @property
    def num_attention_heads(self) -> int:
        return self.dims.num_attention_heads

    @property
    def num_key_value_heads(self) -> int:
        return self.dims.num_key_value_heads

    @property
    def num_hidden_layers(self) -> int:
        return self.dims.num_hidden_layers

    @property
    def hidden_size(self) -> int:
        return self.dims.hidden_size

    @property
    def head_dim(self) -> int:
        return self.dims.head_dim

    @property
    def depth_num_attention_heads(self) -> int:
        return self.dims.depth_num_attention_heads

    @property
    def depth_num_key_value_heads(self) -> int:
        return self.dims.depth_num_key_value_heads

    @property
    def depth_num_hidden_layers(self) -> int:
        return self.dims.depth_num_hidden_layers

    @property
    def depth_hidden_size(self) -> int:
        return self.dims.depth_hidden_size

    @property
    def depth_head_dim(self) -> int:
        return self.dims.depth_head_dim

    @property
    def vocab_size(self) -> int:
        return self.dims.vocab_size

    @property
    def depth_vocab_size(self) -> int:
        return self.dims.vocab_size

    @property
    def max_tokens(self) -> int:
        if self.default_sampling_config.max_tokens is not None:
            return self.default_sampling_config.max_tokens
        return self._max_tokens if self._max_tokens is not None else 1200

    def is_stop_id(self, token_ids: List[int]) -> bool:
        return token_ids[-2] == self.stop_token_id      # the last audio codebook column (csm.py:606-608)

    # ---- prompt side ------------------------------------------------------------------------------------
    def preprocess(self, prompt=None, audio_path: str = None, speaker=0, context=None) -> PreprocessOutput:
        assert audio_path is None
        if not (isinstance(prompt, (tuple, list)) and len(prompt) == 2):
            raise VoxB200Error("no tokenizer / Mimi encoder offline: pass prompt=(ids [T, n_codebooks], masks [T, n_codebooks])")
        ids = torch.as_tensor(prompt[0], dtype=torch.int64)
        masks = torch.as_tensor(prompt[1], dtype=torch.bool)
        assert ids.shape == masks.shape and ids.shape[1] == self.n_codebooks
        return PreprocessOutput(input_tokens=ids, input_masks=masks, repetition_cache=None)

    # ---- engine -------------------------------------------------------------------------------------------
    def engine_for(self, kv_cache: torch.Tensor, page_size: Optional[int] = None) -> CsmEngine:
        key = (kv_cache.data_ptr(), tuple(kv_cache.shape))
        e = self._engines.get(key)
        if e is None:
            e = self._engines[key] = CsmEngine(self.weights, kv_cache, page_size or kv_cache.shape[3],
                                               max_batch=self.max_batch, max_rows=self.max_rows)
        return e

    def frame_device(self, kv_cache: torch.Tensor, attn_wrapper, position_ids: torch.Tensor, n_req: int,
                     input_ids: Optional[torch.Tensor] = None, input_masks: Optional[torch.Tensor] = None,
                     last_rows: Optional[torch.Tensor] = None, sampling_params: Optional[SamplingConfig] = None):
        """One whole frame for ``n_req`` requests on the device -> ids [n_req, n_codebooks] int64 (text column =
        codebook 0, csm.py:693).  Decode: ``input_ids`` None feeds back the frame left by the previous call; pass
        ``input_ids`` [n_req, n_codebooks] to overwrite it first.  Prefill: ``input_ids`` / ``input_masks`` are the
        concatenated prompt rows and ``last_rows`` (int32 [n_req]) every request's last row.  ``attn_wrapper`` must
        have been planned; no host synchronisation happens here."""
        eng = self.engine_for(kv_cache, attn_wrapper.page_size)
        cfg = sampling_params or self.default_sampling_config
        N = self.dims.num_codebooks
        if last_rows is not None:
            return eng.prefill_frame(input_ids, input_masks, position_ids, last_rows, attn_wrapper.plan_rows, cfg)
        if input_ids is not None:
            eng.frame[:N, :n_req].copy_(input_ids[:, :N].t())
        return eng.decode_frame(n_req, position_ids, attn_wrapper.plan_rows, cfg)

    # ---- the reference's step-at-a-time surface (csm.py:637-769) ---------------------------------------------
    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, attn_wrapper, kv_cache: torch.Tensor,
                input_masks: torch.Tensor = None, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (backbone logits [T, 1, vocab], backbone last hidden [T, H])"""
        eng = self.engine_for(kv_cache, attn_wrapper.page_size)
        T = input_ids.shape[0]
        eng.embed_prompt(input_ids.to(torch.int64).contiguous(), input_masks.contiguous())
        logits, hidden = eng.bb.forward(None, position_ids, T, plan=attn_wrapper.plan_rows, want_hidden=True)
        return logits[:, None, :], hidden

    def sampling(self, logits: torch.Tensor, hidden_states: torch.Tensor, requests: List[Request],
                 sampling_params: Optional[SamplingConfig] = None, repetition_cache: Optional[torch.Tensor] = None,
                 cfg_scale: Optional[float] = None, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """Codebook 0 from the backbone logits; returns (ids [B, n_codebooks] with codebook 0 repeated, the depth decoder's
        2-row input [B, 2, H]) and updates the requests like csm.py:665-725."""
        cfg = sampling_params or self.default_sampling_config
        assert logits.shape[1] == 1
        if repetition_cache is not None:
            logits = Sampler.apply_repetition_penalty(logits, repetition_cache, cfg.repetition_penalty)
        ids = Sampler.run_sampling(logits.reshape(-1, self.vocab_size), cfg).view(-1, 1).repeat(1, self.n_codebooks)
        if repetition_cache is not None:
            Sampler.update_repetition_penalty_cache(repetition_cache, ids, cfg.repetition_window)
        c0 = torch.empty(ids.shape[0], self.hidden_size, dtype=BF16, device=ids.device)
        ops.multi_embed_sum(c0, ids[:, 0:1].contiguous(), self.weights.embed_audio, col_offset=self.vocab_size, col0=0)
        hidden_for_depth = torch.stack((hidden_states, c0), dim=1)
        host = ids.cpu()
        for i, req in enumerate(requests):
            req.input_tokens = torch.zeros(1, self.n_codebooks, dtype=torch.long)
            req.input_tokens[0, 0] = host[i, 0]
            req.input_masks = torch.ones(1, self.n_codebooks, dtype=torch.bool)
            req.input_masks[:, -1] = False
            req.lm_output_tokens.append(host[i:i + 1].clone())
            if not self.is_stop_id(host[i].tolist()):
                req.lm_output_audio_tokens.append(req.lm_output_tokens[-1])
            elif req.next_position_id > self.max_tokens:
                req.done_lm_generation, req.finish_reason = True, "max_tokens_reached"
            else:
                req.done_lm_generation, req.finish_reason = True, "stop_id_encountered"
        return ids, hidden_for_depth

    def depth_forward(self, hidden_states: torch.Tensor, position_ids: torch.Tensor, attn_wrapper, kv_cache: torch.Tensor,
                      **kwargs) -> torch.Tensor:
        """rows [R, H] (backbone width) at depth positions -> logits [R, vocab] through the head of the rows' (shared)
        position (csm.py:727-747; the worker never mixes positions in one call apart from the [0, 1] prefill pairs)."""
        eng = next((e for e in self._engines.values() if e.depth_kv.data_ptr() == kv_cache.data_ptr()), None)
        if eng is None:
            raise VoxB200Error("depth_forward runs on the engine's own per-frame cache: pass engine_for(...).depth_kv")
        R = hidden_states.shape[0]
        x = hidden_states.reshape(R, -1).contiguous()
        ops.gemm(x, self.weights.projector, mode=0, out=eng.dp.hidden[:R])
        p = int(position_ids.max())
        head = self.weights.depth.heads[(p - 1) % len(self.weights.depth.heads)]
        return eng.dp.forward(None, position_ids, R, plan=attn_wrapper.plan_rows, head=head)

    def depth_sampling(self, logits: torch.Tensor, i_iteration: int, requests: List[Request],
                       sampling_params: Optional[SamplingConfig] = None, cfg_scale: Optional[float] = None, **kwargs):
        cfg = sampling_params or self.default_sampling_config
        ids = Sampler.run_sampling(logits, cfg)
        emb = torch.empty(ids.shape[0], self.hidden_size, dtype=BF16, device=ids.device)
        ops.multi_embed_sum(emb, ids.view(-1, 1), self.weights.embed_audio, col_offset=self.vocab_size, col0=i_iteration)
        host = ids.cpu().tolist()
        for i, req in enumerate(requests):
            req.input_tokens[0, i_iteration] = host[i]
            req.lm_output_tokens[-1][0, i_iteration] = host[i]
        return ids, emb

    def postprocess(self, token_ids: torch.Tensor, **kwargs) -> torch.Tensor:
        """[B, interval, 33] frame rows -> [B, 1, interval * 1920] (csm.py:771-785): the text column is dropped, codes
        are clamped to the Mimi tables and every chunk is decoded on its own (no state between chunks)."""
        codes = token_ids[:, :, :-1].transpose(1, 2).clamp(0, self.audio_decoder.cfg.bins - 1)
        return self.audio_decoder.decode(codes)
