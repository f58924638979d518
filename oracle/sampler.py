"""Oracle restatement of ``vox_serve/sampling.py`` (reference @6f6b469).

* dispatch order and greedy rule: ``sampling.py:84-118``
* repetition penalty: ``sampling.py:120-146``
* repetition-cache update, including the batch-union quirk (every row is marked with every row's
  sampled id): ``sampling.py:148-178``
* stochastic strategies call FlashInfer's sorting-free rejection sampler (``sampling.py:34,47,59,74``;
  flashinfer-python, third-party, Philox stream drawn from torch's CUDA generator).  Bit-equal
  random streams are not a goal of the north star; the oracle restates the *distribution* each
  strategy samples from (``filtered_probs``) so GPU draws can be checked for support + frequency.

Test infrastructure only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass
class SamplingConfig:  # field-for-field the reference dataclass, sampling.py:8-18
    top_p: Optional[float] = None
    top_k: Optional[int] = None
    min_p: Optional[float] = None
    temperature: float = 1.0
    max_tokens: Optional[int] = None
    repetition_penalty: Optional[float] = None
    repetition_window: Optional[int] = None
    cfg_scale: Optional[float] = None
    greedy: bool = False


def strategy(cfg) -> str:
    """Which branch ``Sampler.run_sampling`` takes (sampling.py:97-118)."""
    if cfg.greedy or cfg.temperature == 0.0:
        return "greedy"
    if cfg.top_k is not None and cfg.top_p is not None:
        return "top_k_top_p"
    if cfg.top_k is not None:
        return "top_k"
    if cfg.top_p is not None:
        return "top_p"
    if cfg.min_p is not None:
        return "min_p"
    return "greedy"


def greedy(logits: torch.Tensor) -> torch.Tensor:
    return torch.argmax(logits, dim=-1)  # first maximal index; int64 (sampling.py:26)


def apply_repetition_penalty(logits: torch.Tensor, cache: torch.Tensor, penalty: float) -> torch.Tensor:
    """logits [B, n_cb, V]; cache [B, W, n_cb, V] bool (sampling.py:137-146)."""
    seen = cache.any(dim=1)
    if logits.shape[1] == 1 and seen.shape[1] != 1:
        seen = seen[:, :1, :]
    out = torch.where((logits > 0) & seen, logits / penalty, logits)
    out = torch.where((out <= 0) & seen, out * penalty, out)
    return out


def update_repetition_cache(cache: torch.Tensor, output_ids: torch.Tensor, window: int) -> None:
    """In place (sampling.py:164-178).  ``cache[..., output_ids] = True`` with a [B, n_cb] index
    marks, for every row, the ids of ALL rows (advanced-index broadcast) -- reproduced as is."""
    ids = output_ids.long()
    if window > 1:
        cache[:, :-1] = cache[:, 1:].clone()
        cache[:, -1].zero_()
        if ids.shape[1] == 1 and cache.shape[2] != 1:
            cache[:, -1, 0, ids[:, 0]] = True
        else:
            cache[:, -1, :, ids] = True
    elif ids.shape[1] == 1 and cache.shape[2] != 1:
        cache[:, :, 0, ids[:, 0]] = True
    else:
        cache[:, :, :, ids] = True


def filtered_probs(logits: torch.Tensor, cfg) -> torch.Tensor:
    """The categorical distribution each stochastic branch draws from, fp32 [N, V].

    Rounding points follow the reference: ``logits / temperature`` and ``softmax`` are taken in the
    logits dtype (bf16 on the hot path, sampling.py:31-32, 44-45, 56, 70-71) before FlashInfer casts
    to fp32.  Filters: top-k keeps the k largest probabilities (ties at the k-th value all kept, as
    the pivot-based kernel does), top-p keeps the smallest prefix of the descending order whose mass
    reaches p, min-p keeps p_i >= min_p * p_max; then renormalise.  ``top_k_top_p`` applies top-k as
    a logit mask first, then softmax, then top-p (FlashInfer default ``top_k_first``).
    """
    kind = strategy(cfg)
    x = logits / cfg.temperature
    if kind == "top_k_top_p":
        xf = x.float()
        kth = torch.topk(xf, min(cfg.top_k, xf.shape[-1]), dim=-1).values[..., -1:]
        xf = xf.masked_fill(xf < kth, float("-inf"))
        p = torch.softmax(xf, dim=-1)
        return _top_p(p, cfg.top_p)
    p = torch.softmax(x, dim=-1).float()
    if kind == "top_k":
        kth = torch.topk(p, min(cfg.top_k, p.shape[-1]), dim=-1).values[..., -1:]
        p = torch.where(p >= kth, p, torch.zeros_like(p))
    elif kind == "top_p":
        return _top_p(p, cfg.top_p)
    elif kind == "min_p":
        p = torch.where(p >= cfg.min_p * p.max(dim=-1, keepdim=True).values, p, torch.zeros_like(p))
    return p / p.sum(dim=-1, keepdim=True)


def _top_p(p: torch.Tensor, top_p: float) -> torch.Tensor:
    sp, idx = torch.sort(p, dim=-1, descending=True)
    csum = torch.cumsum(sp, dim=-1)
    keep_sorted = (csum - sp) < top_p  # keep while the mass BEFORE this item is still < p
    keep = torch.zeros_like(p, dtype=torch.bool).scatter(-1, idx, keep_sorted)
    out = torch.where(keep, p, torch.zeros_like(p))
    return out / out.sum(dim=-1, keepdim=True)
