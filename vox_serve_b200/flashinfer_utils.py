"""Drop-in for ``vox_serve/flashinfer_utils.py``: the names the reference adapters import
(``FlashInferPrefillWrapper``, ``FlashInferDecodeWrapper``, ``FlashInferWrapper``, ``rms_norm``,
``apply_rope_pos_ids``) with the same call signatures, backed by the sm_100a kernels instead of FlashInfer.

``plan()`` still accepts the CPU int32 tensors the reference worker builds (flashinfer_utils.py:60-66,
189-196) but only uploads them and launches the device-side plan kernel: no host work partitioning, no
synchronisation.  ``run`` / ``set_kv_cache`` take one layer's cache ``[pages, 2, page, Hkv, D]`` exactly as
the adapters index it (``kv_cache[i]``, orpheus.py:175-181).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import torch

from . import ops
from ._lib import VoxB200Error


class _PagedWrapper:
    is_prefill = False

    def __init__(self, attn_buffer: Optional[torch.Tensor], n_qo_head: int, n_kv_head: int, n_state: int,
                 page_size: int, batch_size: int = None, max_seq_len: int = None,
                 device: torch.device = torch.device("cuda"), use_cuda_graph: bool = False,
                 max_pages: int = 2048, max_chunks: Optional[int] = None, **_buffers):
        self.device = torch.device(device)
        self.n_qo_head, self.n_kv_head, self.n_state = n_qo_head, n_kv_head, n_state
        self.head_dim = n_state // n_qo_head
        self.page_size = page_size
        self.chunk = ops.attn_chunk_tokens(page_size, n_kv_head)
        self.use_cuda_graph = use_cuda_graph
        self.batch_size = batch_size
        self.max_rows = (max_seq_len or 1024) if self.is_prefill else (batch_size or 64)
        self.max_req = batch_size or 64
        self.max_pages = max_pages
        dev = self.device
        self.d_qo = torch.zeros(self.max_req + 1, dtype=torch.int32, device=dev)
        self.d_indptr = torch.zeros(self.max_req + 1, dtype=torch.int32, device=dev)
        self.d_indices = torch.zeros(max_pages, dtype=torch.int32, device=dev)
        self.d_last = torch.zeros(self.max_req, dtype=torch.int32, device=dev)
        self.plan_rows = ops.RowPlan(self.max_rows, dev)
        self.workspace = None
        self.n_rows = 0
        self.qo_indptr = None

    # -- shared ---------------------------------------------------------------------------------
    def plan_device(self, d_qo: Optional[torch.Tensor], d_indptr: torch.Tensor, d_indices: torch.Tensor,
                    d_last: Optional[torch.Tensor], n_req: int, n_rows: int, kv_len: Optional[torch.Tensor] = None):
        """plan() for page tables that already live on the device (the worker's packed staging buffer): one
        launch, no upload, CUDA-graph capturable."""
        self.n_rows = self.n_rows_padded = n_rows
        self.batch_size = n_req
        ops.plan_rows(self.plan_rows, d_qo, d_indptr, d_indices, d_last, n_req, n_rows, self.page_size, self.chunk,
                      kv_len=kv_len)

    def _upload(self, dst: torch.Tensor, src) -> torch.Tensor:
        t = src if isinstance(src, torch.Tensor) else torch.tensor(src, dtype=torch.int32)
        n = t.numel()
        if n > dst.numel():
            raise VoxB200Error("page table larger than the wrapper's static buffers")
        dst[:n].copy_(t.to(torch.int32), non_blocking=True)
        return dst[:n]

    def set_kv_cache(self, kv_cache: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> None:
        """kv_cache[page, 0, slot] = k ; kv_cache[page, 1, slot] = v (flashinfer_utils.py:144-145, 243-244)."""
        ops.kv_append(kv_cache, k, v, self.plan_rows, n_rows=min(k.shape[0], self.n_rows_padded))

    def run(self, q: torch.Tensor, kv_cache: torch.Tensor) -> torch.Tensor:
        R = q.shape[0]
        if self.workspace is None:
            self.workspace = ops.AttnWorkspace(self.max_rows, self.n_qo_head, self.n_kv_head, self.head_dim,
                                               self.device)
        return ops.paged_attn(q.contiguous(), kv_cache, 0, self.plan_rows, R, self.n_kv_head,
                              self.page_size, self.chunk, self.workspace)


class FlashInferPrefillWrapper(_PagedWrapper):
    is_prefill = True

    def plan(self, qo_indptr, paged_kv_indptr, paged_kv_indices, paged_kv_last_page_len,
             dtype: torch.dtype = torch.bfloat16, n_rows_padded: Optional[int] = None):
        n_req = len(paged_kv_last_page_len)
        qo = self._upload(self.d_qo, qo_indptr)
        ip = self._upload(self.d_indptr, paged_kv_indptr)
        ix = self._upload(self.d_indices, paged_kv_indices)
        la = self._upload(self.d_last, paged_kv_last_page_len)
        total = int(qo_indptr[-1])
        self.n_rows = total
        self.n_rows_padded = total if n_rows_padded is None else n_rows_padded
        self.qo_indptr = qo_indptr if isinstance(qo_indptr, torch.Tensor) else torch.tensor(qo_indptr, dtype=torch.int32)
        self.paged_kv_indptr, self.paged_kv_indices = paged_kv_indptr, paged_kv_indices
        self.paged_kv_last_page_len = paged_kv_last_page_len
        ops.plan_rows(self.plan_rows, qo, ip, ix, la, n_req, self.n_rows_padded, self.page_size, self.chunk)


class FlashInferDecodeWrapper(_PagedWrapper):
    def plan(self, paged_kv_indptr, paged_kv_indices, paged_kv_last_page_len, dtype: torch.dtype = torch.bfloat16):
        n_req = len(paged_kv_last_page_len)
        ip = self._upload(self.d_indptr, paged_kv_indptr)
        ix = self._upload(self.d_indices, paged_kv_indices)
        la = self._upload(self.d_last, paged_kv_last_page_len)
        self.batch_size = n_req
        self.n_rows = self.n_rows_padded = n_req
        self.paged_kv_indptr, self.paged_kv_indices = paged_kv_indptr, paged_kv_indices
        self.paged_kv_last_page_len = paged_kv_last_page_len
        ops.plan_rows(self.plan_rows, None, ip, ix, la, n_req, n_req, self.page_size, self.chunk)


FlashInferWrapper = Union[FlashInferPrefillWrapper, FlashInferDecodeWrapper]


def rms_norm(hidden_states: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return ops.rmsnorm(hidden_states, weight, eps)


def apply_rope_pos_ids(query_states: torch.Tensor, key_states: torch.Tensor, position_ids: torch.Tensor,
                       rope_scale: float = 1.0, rope_theta: float = 10000.0, interleave: bool = False, **kwargs):
    """Same keyword surface as flashinfer_utils.py:270-324 (llama-3.1 kwargs select the smoothed table;
    ``rotary_dim`` restricts the rotation to the leading dims)."""
    d = query_states.shape[-1]
    rd = int(kwargs.get("rotary_dim") or d)
    freq = ops.rope_freq_table(rd, rope_scale, rope_theta, interleave, kwargs.get("low_freq_factor"),
                               kwargs.get("high_freq_factor"), kwargs.get("old_context_len"),
                               device=query_states.device)
    q = query_states.reshape(-1, query_states.shape[-2], d)
    k = key_states.reshape(-1, key_states.shape[-2], d)
    qo, ko = ops.rope(q, k, position_ids.to(torch.int32).reshape(-1), freq, interleave=interleave)
    return qo.view(query_states.shape), ko.view(key_states.shape)
