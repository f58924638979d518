"""Drop-in for ``vox_serve/sampling.py``: same ``SamplingConfig`` fields, same ``Sampler`` classmethods and
dispatch order (sampling.py:84-118), executed by the fused CUDA sampler (csrc/sampler.cu) -- no FlashInfer,
no torch.compile/Triton.

Differences a caller can observe, both inherent to replacing the RNG consumer:
  * stochastic strategies draw from the same filtered distribution as the reference but not the same
    random stream (FlashInfer's rejection sampler consumes torch's Philox state differently);
  * ids are int64 for every strategy (the reference returns int32 from FlashInfer, int64 from argmax).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import Optional

import torch

from . import ops


@dataclass
class SamplingConfig:
    top_p: Optional[float] = None
    top_k: Optional[int] = None
    min_p: Optional[float] = None
    temperature: float = 1.0
    max_tokens: Optional[int] = None
    repetition_penalty: Optional[float] = None
    repetition_window: Optional[int] = None  # -1 for global window
    cfg_scale: Optional[float] = None
    greedy: bool = False


def strategy_of(config: SamplingConfig) -> str:
    """Branch taken by Sampler.run_sampling (sampling.py:97-118)."""
    if config.greedy or config.temperature == 0.0:
        return "greedy"
    if config.top_k is not None and config.top_p is not None:
        return "top_k_top_p"
    if config.top_k is not None:
        return "top_k"
    if config.top_p is not None:
        return "top_p"
    if config.min_p is not None:
        return "min_p"
    return "greedy"


_offset = itertools.count()


class Sampler:
    @classmethod
    def run_sampling(cls, logits: torch.Tensor, config: SamplingConfig, mask_token: int = -1) -> torch.Tensor:
        """logits [N, V] bf16 -> ids [N] int64."""
        kind = strategy_of(config)
        seed = torch.cuda.default_generators[logits.device.index or 0].initial_seed() if logits.is_cuda else 0
        return ops.sample(logits, kind, top_k=config.top_k or 0, top_p=1.0 if config.top_p is None else config.top_p,
                          min_p=config.min_p or 0.0, temperature=config.temperature if kind != "greedy" else 1.0,
                          seed=seed, offset=next(_offset), mask_token=mask_token)

    @classmethod
    def apply_repetition_penalty(cls, logits: torch.Tensor, repetition_cache: torch.Tensor, penalty: float):
        """logits [B, n_cb, V], cache [B, W, n_cb, V] bool -> penalised logits (sampling.py:120-146)."""
        return ops.apply_repetition_penalty(logits, repetition_cache, penalty)

    @classmethod
    def update_repetition_penalty_cache(cls, repetition_cache: torch.Tensor, output_ids: torch.Tensor,
                                        window_size: int, cache_rows: Optional[torch.Tensor] = None) -> None:
        """In place, including the reference's batch-union marking (sampling.py:148-178).  ``cache_rows`` maps
        batch rows onto rows of a larger slot-resident cache."""
        ops.update_repetition_cache(repetition_cache, output_ids, window_size, cache_rows=cache_rows)

    @classmethod
    def sample_fused(cls, logits: torch.Tensor, config: SamplingConfig, repetition_cache: Optional[torch.Tensor],
                     mask_token: int = -1, rng_state: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None, cache_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
        """penalty + strategy + draw in one call: logits [B, n_cb, V] -> ids [B, n_cb] int64.
        (what orpheus.py:431-438 does in three steps)."""
        B, C, V = logits.shape
        kind = strategy_of(config)
        seed = torch.cuda.default_generators[logits.device.index or 0].initial_seed()
        ids = ops.sample(logits.view(B * C, V), kind, rep_cache=repetition_cache,
                         penalty=config.repetition_penalty or 1.0, logit_codebooks=C, top_k=config.top_k or 0,
                         top_p=1.0 if config.top_p is None else config.top_p, min_p=config.min_p or 0.0,
                         temperature=config.temperature if kind != "greedy" else 1.0, seed=seed,
                         offset=next(_offset) if rng_state is None else 0, mask_token=mask_token,
                         rng_state=rng_state, out=out, cache_rows=cache_rows)
        return ids.view(B, C)
