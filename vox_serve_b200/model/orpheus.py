"""Orpheus-3B adapter on the sm_100a kernels: same class name, constructor keywords, properties and
``preprocess / forward / sampling / postprocess`` contract as ``vox_serve/model/orpheus.py:224-507``.

What differs from the reference adapter is only *where the arithmetic runs*:
  * ``forward`` drives ``LlamaEngine`` (fused C-ABI launches, engine.py) instead of an ``nn.Module`` graph;
  * ``sampling`` is one fused penalty + filter + draw launch plus the cache update (csrc/sampler.cu) and keeps
    the reference's host-side ``update_req_states`` coroutine (orpheus.py:449-473);
  * ``postprocess`` de-interleaves the 7-token frames on the device and runs the SNAC kernels, computing only
    the ``[2048:4096]`` samples the reference keeps (orpheus.py:506).

Weights: a local HF directory (``config.json`` + ``*.safetensors``), an explicit ``state_dict=``, or
``model_name="orpheus-synthetic[:seed]"`` (seeded N(0, 0.02) weights at the true Orpheus shapes -- the image
has no network, so this is what bench.py and the parity tests use).
"""
from __future__ import annotations

import glob
import json
import math
import os
from typing import Any, Dict, List, Optional

import torch

from .. import ops
from .._lib import VoxB200Error
from ..engine import LlamaDims, LlamaEngine, LlamaWeights, hf_layer_names
from ..requests import Request
from ..sampling import Sampler, SamplingConfig
from ..tokenizer.snac import SNAC
from .base import BaseLM, PreprocessOutput

BF16 = torch.bfloat16


def synthetic_state_dict(dims: LlamaDims, seed: int = 0, device="cuda", lm_head_scale: float = 8.0) -> Dict[str, torch.Tensor]:
    """Seeded bf16 weights under HF Llama names, generated on ``device`` (3.3 B parameters take ~1 s on a B200).
    lm_head is scaled up so greedy top-1/top-2 margins sit well above bf16 noise."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)

    def rnd(*shape, std=0.02, mean=0.0):
        # generated in row chunks to bound the fp32 temporary for the 157 k x 3072 matrices
        out = torch.empty(*shape, dtype=BF16, device=dev)
        flat = out.view(shape[0], -1) if len(shape) > 1 else out.view(-1, 1)
        step = max(1, (1 << 26) // max(1, flat.shape[1]))
        for r in range(0, flat.shape[0], step):
            blk = flat[r:r + step]
            blk.copy_(torch.randn(blk.shape, generator=g, dtype=torch.float32, device=dev) * std + mean)
        return out

    H, I = dims.hidden_size, dims.intermediate_size
    hq, hkv = dims.num_attention_heads * dims.head_dim, dims.num_key_value_heads * dims.head_dim
    sd = {"model.embed_tokens.weight": rnd(dims.vocab_size, H, std=1.0)}
    for i in range(dims.num_hidden_layers):
        n = hf_layer_names(i)
        sd[n["ln1"]], sd[n["ln2"]] = rnd(H, std=0.1, mean=1.0), rnd(H, std=0.1, mean=1.0)
        sd[n["q"]], sd[n["k"]], sd[n["v"]], sd[n["o"]] = rnd(hq, H), rnd(hkv, H), rnd(hkv, H), rnd(H, hq)
        sd[n["gate"]], sd[n["up"]], sd[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
    sd["model.norm.weight"] = rnd(H, std=0.1, mean=1.0)
    sd["lm_head.weight"] = rnd(dims.vocab_size, H, std=0.02 * lm_head_scale)
    return sd


def _load_hf_dir(path: str) -> (LlamaDims, Dict[str, torch.Tensor]):
    from safetensors.torch import load_file

    with open(os.path.join(path, "config.json")) as f:
        c = json.load(f)
    rs = c.get("rope_scaling") or {}
    dims = LlamaDims(c["hidden_size"], c["num_hidden_layers"], c["num_attention_heads"], c["num_key_value_heads"],
                     c.get("head_dim", c["hidden_size"] // c["num_attention_heads"]), c["intermediate_size"],
                     c["vocab_size"], c.get("rms_norm_eps", 1e-5), c.get("rope_theta", 500000.0),
                     rs.get("factor", 32.0), rs.get("low_freq_factor", 1.0), rs.get("high_freq_factor", 4.0),
                     rs.get("original_max_position_embeddings", 8192))
    sd: Dict[str, torch.Tensor] = {}
    for fn in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
        sd.update(load_file(fn))
    if not sd:
        raise VoxB200Error(f"no *.safetensors under {path}")
    return dims, sd


class OrpheusModel(BaseLM):
    STOP_TOKEN_ID = 128258           # orpheus.py:258
    AUDIO_ID_BASE = 128256 + 10      # orpheus.py:479-481

    def __init__(self, model_name, dtype=BF16, device="cuda:0", tokenizer_path="canopylabs/orpheus-3b-0.1-ft",
                 enable_torch_compile=False, audio_decoder_device=None, state_dict: Optional[Dict] = None,
                 dims: Optional[LlamaDims] = None, snac: Optional[SNAC] = None, snac_state_dict: Optional[Dict] = None,
                 stop_token_id: Optional[int] = None, audio_id_base: Optional[int] = None,
                 max_tokens: Optional[int] = None, mask_stop_token: bool = False):
        if model_name == "orpheus":
            model_name = "canopylabs/orpheus-3b-0.1-ft"
        if dtype != BF16:
            raise VoxB200Error("the B200 decode path computes in bf16 only")
        super().__init__(model_name, device, dtype, enable_torch_compile, audio_decoder_device)
        if not torch.cuda.is_available():
            raise VoxB200Error("OrpheusModel needs a CUDA device: there is no CPU path")
        seed = 0
        if state_dict is None:
            if os.path.isdir(model_name):
                dims, state_dict = _load_hf_dir(model_name)
            elif model_name.startswith("orpheus-synthetic"):
                # "orpheus-synthetic[-tiny][:seed]": -tiny = a 2-layer, hidden-768 model with the true vocabulary and
                # a 64-channel SNAC (integration tests that construct the worker by NAME, like the reference's
                # Scheduler does: scheduler/base.py:63-91)
                seed = int(model_name.split(":")[1]) if ":" in model_name else 0
                tiny = model_name.split(":")[0].endswith("-tiny")
                dims = dims or (LlamaDims(768, 2, 6, 2, 128, 1024, 156940) if tiny else LlamaDims.orpheus_3b())
                state_dict = synthetic_state_dict(dims, seed, device)
                if tiny and snac is None:
                    snac = SNAC(encoder_dim=4, encoder_rates=(2, 2, 2, 2), decoder_dim=64, device=audio_decoder_device or device)
                    snac.load_state_dict(snac.synthetic_state_dict(seed=seed + 1))
            else:
                raise VoxB200Error(
                    f"'{model_name}' is not a local directory: this build has no network access. Pass a local HF "
                    "directory, state_dict=..., or model_name='orpheus-synthetic[:seed]'")
        self.dims = dims or LlamaDims.orpheus_3b()
        self.weights = LlamaWeights.from_state_dict(state_dict, self.dims, device)
        del state_dict
        self.available_voices = ["tara", "leah", "jess", "leo", "dan", "mia", "zac", "zoe"]
        self.text_tokenizer = self._load_tokenizer(tokenizer_path)
        if snac is None:
            snac = SNAC(device=self.audio_decoder_device)
            snac.load_state_dict(snac_state_dict if snac_state_dict is not None
                                 else snac.synthetic_state_dict(seed=seed + 1))
        self.audio_decoder = snac
        self.stop_token_id = self.STOP_TOKEN_ID if stop_token_id is None else stop_token_id
        self.audio_id_base = self.AUDIO_ID_BASE if audio_id_base is None else audio_id_base
        self._max_tokens = max_tokens
        # benchmark hook ("fixed 700 decode steps/request (stop id masked out)")
        self.mask_stop_token = mask_stop_token
        self.default_sampling_config = SamplingConfig(top_k=None, top_p=0.8, min_p=None, temperature=0.6,
                                                      repetition_penalty=1.3, repetition_window=-1, cfg_scale=None)
        self._engines: Dict[Any, LlamaEngine] = {}
        self.max_rows = 1024 + 64
        # device {seed, offset}: the sampler advances the offset itself, so CUDA-graph replays keep drawing
        seed64 = torch.cuda.default_generators[torch.device(device).index or 0].initial_seed() & ((1 << 63) - 1)
        self.rng_state = torch.tensor([seed64, 0, 0], dtype=torch.int64, device=device)

    # ---- static facts (orpheus.py:276-330) --------------------------------------------------------
    n_codebooks = 1
    detokenize_interval = 28
    detokenize_overlap = 21
    n_channels = 1
    output_audio_length = 2048

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
    def vocab_size(self) -> int:
        return self.dims.vocab_size

    @property
    def max_tokens(self) -> int:
        if self.default_sampling_config.max_tokens is not None:
            return self.default_sampling_config.max_tokens
        return self._max_tokens if self._max_tokens is not None else 1200

    def is_stop_id(self, token_ids: List[int]) -> bool:
        return token_ids[0] == self.stop_token_id

    # ---- prompt side ----------------------------------------------------------------------------
    def _load_tokenizer(self, tokenizer_path):
        if tokenizer_path and os.path.isdir(str(tokenizer_path)):
            from transformers import AutoTokenizer

            return AutoTokenizer.from_pretrained(tokenizer_path)
        return None       # no tokenizer files on disk (and no network): prompts must arrive as token ids

    def _validate_voice(self, voice):
        if voice and voice not in self.available_voices:
            raise ValueError(f"Voice {voice} is not available for model {self.model_name}")

    def _format_prompt(self, prompt, voice="tara", model_type="larger") -> torch.Tensor:
        """``[128259] + ids + [128009, 128260, 128261, 128257]`` (orpheus.py:347-368)."""
        if isinstance(prompt, str):
            if self.text_tokenizer is None:
                raise VoxB200Error("no text tokenizer on disk: pass the prompt as a sequence of token ids")
            text = f"{voice}: {prompt}" if voice else prompt
            ids = self.text_tokenizer(text, return_tensors="pt").input_ids[0]
            if not voice:            # the reference returns the bare tokenizer output in this branch (:366-368)
                return ids.to(torch.int64)
        else:
            ids = torch.as_tensor(prompt, dtype=torch.int64).reshape(-1)
        start = torch.tensor([128259], dtype=torch.int64)
        end = torch.tensor([128009, 128260, 128261, 128257], dtype=torch.int64)
        return torch.cat((start, ids, end))

    def new_repetition_cache(self, device=None) -> Optional[torch.Tensor]:
        c = self.default_sampling_config
        if c.repetition_penalty is None or c.repetition_window is None or c.repetition_penalty == 1.0:
            return None
        return torch.zeros(c.repetition_window if c.repetition_window > 0 else 1, self.n_codebooks, self.vocab_size,
                           dtype=torch.bool, device=device or self.device)

    def preprocess(self, prompt=None, audio_path: str = None, voice="tara", model_type="larger",
                   repetition_cache: Optional[torch.Tensor] = None) -> PreprocessOutput:
        """``repetition_cache`` lets the worker hand in a zeroed slot-resident cache row instead of allocating a
        fresh 157 KB tensor per request (orpheus.py:381-394 allocates)."""
        assert audio_path is None
        self._validate_voice(voice)
        input_ids = self._format_prompt(prompt, voice, model_type).view(-1, 1)
        if repetition_cache is None:
            repetition_cache = self.new_repetition_cache()
        return PreprocessOutput(input_tokens=input_ids, repetition_cache=repetition_cache)

    # ---- LM step ------------------------------------------------------------------------------------
    def engine_for(self, kv_cache: torch.Tensor, page_size: Optional[int] = None) -> LlamaEngine:
        key = (kv_cache.data_ptr(), tuple(kv_cache.shape))
        e = self._engines.get(key)
        if e is None:
            e = LlamaEngine(self.weights, kv_cache, page_size or kv_cache.shape[3], max_rows=self.max_rows,
                            max_seq_len=max(2304, self.max_tokens + 1024))
            self._engines[key] = e
        return e

    def forward(self, input_ids: torch.Tensor, position_ids: torch.Tensor, attn_wrapper, kv_cache: torch.Tensor,
                logits_rows: Optional[torch.Tensor] = None, n_rows: Optional[int] = None,
                logits_rows_minus_one: bool = False, **kwargs) -> torch.Tensor:
        """input_ids [T, 1] (any int dtype) / position_ids [T] int32 -> logits [T, 1, V] bf16 (orpheus.py:398-417).
        ``attn_wrapper`` must have been planned (its device row plan is what the kernels read).
        ``logits_rows`` (int32, device) restricts lm_head to those rows (the worker passes qo_indptr[1:]-1 for
        prefill instead of slicing a [T, V] tensor afterwards, cuda_graph_worker.py:900-902)."""
        eng = self.engine_for(kv_cache, attn_wrapper.page_size)
        ids = input_ids.reshape(-1) if input_ids.dim() == 1 else input_ids[:, 0]
        if ids.dtype != torch.int32 or not ids.is_contiguous():
            ids = ids.to(torch.int32).contiguous()
        T = ids.numel() if n_rows is None else n_rows
        logits = eng.forward(ids, position_ids, T, last_rows=logits_rows, plan=attn_wrapper.plan_rows,
                             last_rows_offset=-1 if logits_rows_minus_one else 0)
        return logits[:, None, :]

    def sampling_device(self, logits: torch.Tensor, sampling_params: Optional[SamplingConfig] = None,
                        repetition_cache: Optional[torch.Tensor] = None, cache_rows: Optional[torch.Tensor] = None,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device half of ``sampling``: penalty -> strategy -> draw -> cache update (orpheus.py:431-445).
        No host synchronisation, no allocation when ``out`` is given: CUDA-graph capturable."""
        cfg = sampling_params or self.default_sampling_config
        ids = Sampler.sample_fused(logits, cfg, repetition_cache, cache_rows=cache_rows, rng_state=self.rng_state,
                                   mask_token=self.stop_token_id if self.mask_stop_token else -1, out=out)
        if repetition_cache is not None:
            Sampler.update_repetition_penalty_cache(repetition_cache, ids, cfg.repetition_window, cache_rows=cache_rows)
        return ids

    def sampling_host(self, output_ids: torch.Tensor, requests: List[Request],
                      repetition_cache: Optional[torch.Tensor] = None, ids_host=None, ready=None,
                      cache_rows_host: Optional[List[int]] = None):
        """Host half: ``req.input_tokens`` + the ``update_req_states`` coroutine (orpheus.py:447-475).
        ``ids_host`` (pinned int64 [B, 1]) + ``ready`` (CUDA event) are the asynchronous D2H copy of
        ``output_ids`` started by the worker; without them the coroutine copies synchronously."""
        for i, req in enumerate(requests):
            req.input_tokens = output_ids[i:i + 1]
        stop_id, max_tokens = self.stop_token_id, self.max_tokens

        async def update_req_states():
            if ids_host is not None:
                ready.synchronize()
                host = ids_host[:len(requests)].clone()
            else:
                host = output_ids.to("cpu")
            col0 = host[:, 0].tolist()
            for i, req in enumerate(requests):
                tok = host[i:i + 1]
                req.lm_output_tokens.append(tok)
                req.lm_output_audio_tokens.append(tok)
                if col0[i] == stop_id:
                    req.lm_output_audio_tokens.pop()
                    req.done_lm_generation = True
                    req.finish_reason = "stop_id_encountered"
            for req in requests:
                if req.next_position_id > max_tokens:
                    req.done_lm_generation = True
                    req.finish_reason = "max_tokens_reached"
            if repetition_cache is not None:
                for i, req in enumerate(requests):
                    row = i if cache_rows_host is None else cache_rows_host[i]
                    req.repetition_cache = repetition_cache[row]

        return update_req_states()

    def sampling(self, logits: torch.Tensor, requests: List[Request], sampling_params: Optional[SamplingConfig] = None,
                 repetition_cache: Optional[torch.Tensor] = None, cfg_scale: Optional[float] = None, **kwargs):
        output_ids = self.sampling_device(logits, sampling_params, repetition_cache)
        return output_ids, self.sampling_host(output_ids, requests, repetition_cache)

    # ---- vocoder ------------------------------------------------------------------------------------
    def postprocess(self, token_ids: torch.Tensor, noises=None, **kwargs) -> torch.Tensor:
        """[B, 28, 1] LM ids -> [B, 1, 2048] fp32 (orpheus.py:483-507)."""
        c0, c1, c2 = ops.orpheus_window_codes(token_ids.reshape(-1, 28), self.audio_id_base)
        n = c2.shape[1] * self.audio_decoder.vq_strides[-1] * math.prod(self.audio_decoder.decoder_rates)   # 8192
        return self.audio_decoder.decode([c0, c1, c2], noises=noises, out_range=(n // 4, n // 2))
