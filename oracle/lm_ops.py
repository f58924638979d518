"""Oracle restatement of the L1 operators the reference imports from FlashInfer.

Reference call sites: ``vox_serve/flashinfer_utils.py:251-324`` (rms_norm, apply_rope_pos_ids),
``:86-145`` (prefill slot maps + scatter), ``:189-244`` (decode slot map + scatter + run).
Third-party arithmetic restated from the installed FlashInfer headers:
``include/flashinfer/norm.cuh:64-101`` and ``include/flashinfer/pos_enc.cuh:594-617, 129-147,
1399-1400, 1538-1539``.

Plain torch on CPU; fp32 internal math, outputs rounded once to the input dtype (bf16 on the hot
path).  Test infrastructure only (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch


# --------------------------------------------------------------------------------------------
# RMSNorm  (flashinfer_utils.py:251-267 -> flashinfer.norm.rmsnorm, norm.cuh:64-101)
# --------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    xf = x.float()
    ss = (xf * xf).sum(dim=-1, keepdim=True)
    rcp = torch.rsqrt(ss / x.shape[-1] + eps)
    return ((xf * rcp) * weight.float()).to(x.dtype)


# --------------------------------------------------------------------------------------------
# RoPE at explicit positions (flashinfer_utils.py:270-324, pos_enc.cuh:594-617)
# --------------------------------------------------------------------------------------------
def rope_freqs(
    rotary_dim: int,
    rope_scale: float,
    rope_theta: float,
    interleave: bool,
    low_freq_factor: Optional[float] = None,
    high_freq_factor: Optional[float] = None,
    old_context_len: Optional[float] = None,
) -> torch.Tensor:
    """Per-element frequency vector of length ``rotary_dim`` (fp32).

    Plain variant: smooth_a = smooth_b = 0 (pos_enc.cuh:1399-1400) so ``freq = theta^-e / scale``.
    Llama-3.1 variant (pos_enc.cuh:1538-1539): smooth interpolation between scaled / unscaled.
    """
    i = torch.arange(rotary_dim, dtype=torch.float32)
    if interleave:
        e = 2.0 * torch.floor(i / 2.0) / rotary_dim
    else:
        e = 2.0 * torch.remainder(i, rotary_dim // 2) / rotary_dim
    freq = torch.pow(torch.tensor(1.0 / rope_theta, dtype=torch.float32), e)
    if low_freq_factor is not None or high_freq_factor is not None or old_context_len is not None:
        lo = 1.0 if low_freq_factor is None else float(low_freq_factor)
        hi = 4.0 if high_freq_factor is None else float(high_freq_factor)
        ctx = 8192.0 if old_context_len is None else float(old_context_len)
        smooth_a = ctx / (2 * math.pi * hi - 2 * math.pi * lo)
        smooth_b = -1.0 / (hi / lo - 1.0)
    else:
        smooth_a, smooth_b = 0.0, 0.0
    smooth = torch.clamp(freq * smooth_a + smooth_b, 0.0, 1.0)
    return (1 - smooth) * (freq * (1.0 / rope_scale)) + smooth * freq


def _rope_one(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, rotary_dim: int, interleave: bool):
    """x: [T, H, D]; cos/sin: [T, rotary_dim] fp32."""
    xf = x.float()
    rot = xf[..., :rotary_dim]
    if interleave:
        # pairs (2j, 2j+1): out[2j] = x[2j] c - x[2j+1] s ; out[2j+1] = x[2j+1] c + x[2j] s
        x_even, x_odd = rot[..., 0::2], rot[..., 1::2]
        perm = torch.stack((-x_odd, x_even), dim=-1).flatten(-2)
    else:
        half = rotary_dim // 2
        perm = torch.cat((-rot[..., half:], rot[..., :half]), dim=-1)
    out = rot * cos[:, None, :] + perm * sin[:, None, :]
    res = xf.clone()
    res[..., :rotary_dim] = out
    return res.to(x.dtype)


def apply_rope_pos_ids(
    q: torch.Tensor,
    k: torch.Tensor,
    position_ids: torch.Tensor,
    rope_scale: float = 1.0,
    rope_theta: float = 10000.0,
    interleave: bool = False,
    rotary_dim: Optional[int] = None,
    **llama31,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """q: [T, Hq, D], k: [T, Hkv, D], position_ids: [T] int.  Returns new tensors."""
    d = q.shape[-1]
    rd = d if rotary_dim is None else int(rotary_dim)
    freq = rope_freqs(
        rd, rope_scale, rope_theta, interleave,
        llama31.get("low_freq_factor"), llama31.get("high_freq_factor"), llama31.get("old_context_len"),
    )
    embed = position_ids.to(torch.float32)[:, None] * freq[None, :]
    cos, sin = torch.cos(embed), torch.sin(embed)
    return _rope_one(q, cos, sin, rd, interleave), _rope_one(k, cos, sin, rd, interleave)


# --------------------------------------------------------------------------------------------
# Paged KV bookkeeping + attention (flashinfer_utils.py:60-145, 189-244)
# layer cache layout: [n_pages, 2, page_size, Hkv, D]  ("NHD" pages)
# --------------------------------------------------------------------------------------------
def decode_slots(indptr: Sequence[int], indices: Sequence[int], last_page_len: Sequence[int]):
    """New-token (page, slot) per request: flashinfer_utils.py:217-219."""
    pages = [indices[indptr[i + 1] - 1] for i in range(len(last_page_len))]
    slots = [l - 1 for l in last_page_len]
    return pages, slots


def prefill_slots(
    qo_indptr: Sequence[int], indptr: Sequence[int], indices: Sequence[int],
    last_page_len: Sequence[int], page_size: int,
):
    """Per-token (page, slot) for a ragged prefill batch: flashinfer_utils.py:86-124."""
    pages, slots = [], []
    for r in range(len(last_page_len)):
        n_new = qo_indptr[r + 1] - qo_indptr[r]
        n_pages = indptr[r + 1] - indptr[r]
        kv_len = (n_pages - 1) * page_size + last_page_len[r]
        for j in range(n_new):
            g = kv_len - n_new + j
            pages.append(indices[indptr[r] + g // page_size])
            slots.append(g % page_size)
    return pages, slots


def kv_append(layer_cache: torch.Tensor, k: torch.Tensor, v: torch.Tensor, pages, slots) -> None:
    """kv[page,0,slot] = k ; kv[page,1,slot] = v   (flashinfer_utils.py:144-145, 243-244)."""
    p = torch.as_tensor(pages, dtype=torch.long)
    s = torch.as_tensor(slots, dtype=torch.long)
    layer_cache[p, 0, s] = k
    layer_cache[p, 1, s] = v


def _gather_kv(layer_cache, indices, start, end, last_len, page_size):
    pg = torch.as_tensor(list(indices[start:end]), dtype=torch.long)
    kv_len = (end - start - 1) * page_size + last_len
    blk = layer_cache[pg]  # [n, 2, page, H, D]
    k = blk[:, 0].reshape(-1, blk.shape[-2], blk.shape[-1])[:kv_len]
    v = blk[:, 1].reshape(-1, blk.shape[-2], blk.shape[-1])[:kv_len]
    return k, v, kv_len


def _attend(q, k, v, sm_scale, causal_offset: Optional[int]):
    """q [Tq,Hq,D], k/v [Tk,Hkv,D] -> [Tq,Hq,D]; fp32 softmax; GQA by head grouping."""
    hq, hkv = q.shape[1], k.shape[1]
    g = hq // hkv
    qf = q.float().transpose(0, 1)                                   # [Hq,Tq,D]
    kf = k.float().transpose(0, 1).repeat_interleave(g, dim=0)       # [Hq,Tk,D]
    vf = v.float().transpose(0, 1).repeat_interleave(g, dim=0)
    s = torch.matmul(qf, kf.transpose(1, 2)) * sm_scale              # [Hq,Tq,Tk]
    if causal_offset is not None:
        tq, tk = s.shape[1], s.shape[2]
        qi = torch.arange(tq)[:, None] + causal_offset
        kj = torch.arange(tk)[None, :]
        s = s.masked_fill(kj > qi, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf)                                          # [Hq,Tq,D]
    return o.transpose(0, 1).to(q.dtype)


def paged_attention_decode(
    q: torch.Tensor, layer_cache: torch.Tensor, indptr, indices, last_page_len, page_size: int,
    sm_scale: Optional[float] = None,
) -> torch.Tensor:
    """One query token per request over its pages (flashinfer_utils.py:228-230).

    q: [B, Hq, D] -> [B, Hq, D].  sm_scale defaults to 1/sqrt(D) as FlashInfer does.
    """
    d = q.shape[-1]
    sc = (1.0 / math.sqrt(d)) if sm_scale is None else sm_scale
    out = torch.empty_like(q)
    for r in range(q.shape[0]):
        k, v, _ = _gather_kv(layer_cache, indices, indptr[r], indptr[r + 1], last_page_len[r], page_size)
        out[r : r + 1] = _attend(q[r : r + 1], k, v, sc, None)
    return out


def paged_attention_prefill(
    q: torch.Tensor, layer_cache: torch.Tensor, qo_indptr, indptr, indices, last_page_len,
    page_size: int, sm_scale: Optional[float] = None,
) -> torch.Tensor:
    """Ragged causal prefill over paged KV (flashinfer_utils.py:68-80, 132; causal=True).

    Query token j of request r sits at kv position kv_len - n_new + j and sees keys <= that.
    Rows beyond qo_indptr[-1] (CUDA-graph padding) are returned as zeros.
    """
    d = q.shape[-1]
    sc = (1.0 / math.sqrt(d)) if sm_scale is None else sm_scale
    out = torch.zeros_like(q)
    for r in range(len(last_page_len)):
        a, b = qo_indptr[r], qo_indptr[r + 1]
        if b == a:
            continue
        k, v, kv_len = _gather_kv(layer_cache, indices, indptr[r], indptr[r + 1], last_page_len[r], page_size)
        out[a:b] = _attend(q[a:b], k, v, sc, kv_len - (b - a))
    return out


class PagedWrapperCPU:
    """plan / set_kv_cache / run object with the attribute surface the reference adapters use
    (flashinfer_utils.py:11-244): injected into ``model.forward(..., attn_wrapper=...)``.
    ``mode`` is "decode" or "prefill"."""

    def __init__(self, mode: str, page_size: int):
        self.mode, self.page_size = mode, page_size
        self.qo_indptr = None

    def plan(self, *args):
        if self.mode == "decode":
            indptr, indices, last = [list(map(int, a)) for a in args[:3]]
            self.indptr, self.indices, self.last = indptr, indices, last
            self.pages, self.slots = decode_slots(indptr, indices, last)
        else:
            qo, indptr, indices, last = [list(map(int, a)) for a in args[:4]]
            self.qo_indptr = torch.as_tensor(qo, dtype=torch.int32)
            self.qo, self.indptr, self.indices, self.last = qo, indptr, indices, last
            self.pages, self.slots = prefill_slots(qo, indptr, indices, last, self.page_size)

    def set_kv_cache(self, kv_cache, k, v):
        n = len(self.pages)
        kv_append(kv_cache, k[:n], v[:n], self.pages, self.slots)

    def run(self, q, kv_cache):
        if self.mode == "decode":
            return paged_attention_decode(q, kv_cache, self.indptr, self.indices, self.last, self.page_size)
        return paged_attention_prefill(q, kv_cache, self.qo, self.indptr, self.indices, self.last, self.page_size)
