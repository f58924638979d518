"""Tensor-level host wrappers over the C ABI (include/vb_api.h).

PyTorch is used here only for device memory and streams: every function takes CUDA tensors, passes raw
pointers + sizes to libvoxb200.so on the current stream and returns torch tensors.  There is no fallback
path: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import VoxB200Error, call

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise VoxB200Error("vox_serve_b200 ops need CUDA tensors (there is no CPU path)")


_dev_info: Dict[int, Tuple[int, int]] = {}


def launch_count() -> int:
    """Kernels (and memset nodes) enqueued through the C ABI so far in this process."""
    return _lib.launch_counter[0]


def device_info() -> Tuple[int, int]:
    dev = torch.cuda.current_device()
    if dev not in _dev_info:
        sm, smem = ctypes.c_int(0), ctypes.c_int(0)
        call("vb_device_info", ctypes.byref(sm), ctypes.byref(smem))
        _dev_info[dev] = (sm.value, smem.value)
    return _dev_info[dev]


# ----------------------------------------------------------------------------------------------
# TMA descriptors
# ----------------------------------------------------------------------------------------------
class TensorMap:
    """128-byte CUtensorMap held in 64-byte aligned host memory; keeps the tensor alive."""

    __slots__ = ("_raw", "ptr", "owner")

    def __init__(self, owner: torch.Tensor):
        self._raw = ctypes.create_string_buffer(128 + 64)
        base = ctypes.addressof(self._raw)
        self.ptr = (base + 63) & ~63
        self.owner = owner


def tensor_map_kv(kv_cache: torch.Tensor, box_tokens: int) -> TensorMap:
    """kv_cache: [layers, pages, 2, page_size, n_kv, head_dim] bf16 (or a single layer [pages, 2, ...])."""
    _need_cuda(kv_cache)
    assert kv_cache.dtype == BF16 and kv_cache.is_contiguous()
    if kv_cache.dim() == 6:
        n_slabs = kv_cache.shape[0] * kv_cache.shape[1]
    else:
        assert kv_cache.dim() == 5
        n_slabs = kv_cache.shape[0]
    page_size, n_kv, d = kv_cache.shape[-3], kv_cache.shape[-2], kv_cache.shape[-1]
    tm = TensorMap(kv_cache)
    call("vb_tensor_map_kv", tm.ptr, kv_cache.data_ptr(), n_slabs, page_size, n_kv, d, box_tokens)
    return tm


# tensor maps of the engine's static activation buffers; bounded: ad-hoc callers (tests, the operator shims on
# transient tensors) would otherwise pin every tensor they ever passed
_map2d_cache: "OrderedDict[Tuple, TensorMap]" = None
_MAP2D_CACHE_MAX = 256


def tensor_map_2d(t: torch.Tensor, box_rows: int, cache: bool = True) -> TensorMap:
    _need_cuda(t)
    assert t.dtype == BF16 and t.dim() == 2 and t.stride(1) == 1
    global _map2d_cache
    if _map2d_cache is None:
        from collections import OrderedDict

        _map2d_cache = OrderedDict()
    key = (t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), box_rows)
    if cache and key in _map2d_cache:
        _map2d_cache.move_to_end(key)
        return _map2d_cache[key]
    tm = TensorMap(t)
    call("vb_tensor_map_2d_bf16", tm.ptr, t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), box_rows)
    if cache:
        _map2d_cache[key] = tm
        while len(_map2d_cache) > _MAP2D_CACHE_MAX:
            _map2d_cache.popitem(last=False)
    return tm


# ----------------------------------------------------------------------------------------------
# norm / rope / paging
# ----------------------------------------------------------------------------------------------
class TiledAct:
    """An activation matrix [T, K] in the tiled XT(t_tile) layout of vb_api.h (what a GEMM stage looks like in shared
    memory): producers write it directly (rmsnorm / reduce_residual_rmsnorm / paged_attn / gemm mode 2 with `out=` a
    TiledAct), ops.gemm consumes it with one linear bulk copy per pipeline stage."""
    __slots__ = ("data", "T", "K", "t_tile")

    def __init__(self, T: int, K: int, device, max_T: Optional[int] = None):
        self.T, self.K = T, K
        self.t_tile = gemm_t_tile(T)
        cap = gemm_t_tile(max_T or T)
        blocks = (max(T, max_T or T) + cap - 1) // cap
        self.data = torch.zeros(blocks * ((K + 63) // 64) * cap * 64, dtype=BF16, device=device)

    def view_rows(self, T: int) -> "TiledAct":
        """the same storage re-interpreted for T rows (t_tile follows T)"""
        v = TiledAct.__new__(TiledAct)
        v.data, v.T, v.K, v.t_tile = self.data, T, self.K, gemm_t_tile(T)
        assert ((T + v.t_tile - 1) // v.t_tile) * ((self.K + 63) // 64) * v.t_tile * 64 <= self.data.numel()
        return v

    def to_rows(self) -> torch.Tensor:
        """row-major copy [T, K] (tests / debugging)"""
        t, K, nkb = self.t_tile, self.K, (self.K + 63) // 64
        blocks = (self.T + t - 1) // t
        x = self.data[: blocks * nkb * t * 64].view(blocks, nkb, t, 8, 8)
        r = torch.arange(t, device=x.device) & 7
        c = torch.arange(8, device=x.device)
        src = (c.view(1, 8) ^ r.view(t, 1))                      # logical chunk c of row r sits at chunk c ^ (r & 7)
        x = torch.gather(x, 3, src.view(1, 1, t, 8, 1).expand(blocks, nkb, t, 8, 8))
        return x.permute(0, 2, 1, 3, 4).reshape(blocks * t, nkb * 64)[: self.T, :K].contiguous()


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float, out=None):
    """out: None / a tensor -> row-major result; a TiledAct -> written in the tiled layout (returned as is)."""
    _need_cuda(x, weight)
    assert x.dtype == BF16 and weight.dtype == BF16
    x2 = x.contiguous().view(-1, x.shape[-1])
    if isinstance(out, TiledAct):
        assert out.T == x2.shape[0] and out.K == x2.shape[1]
        call("vb_rmsnorm", out.data.data_ptr(), x2.data_ptr(), weight.data_ptr(), x2.shape[0], x2.shape[1], float(eps),
             out.t_tile, _stream())
        return out
    out = torch.empty_like(x2) if out is None else out
    call("vb_rmsnorm", out.data_ptr(), x2.data_ptr(), weight.data_ptr(), x2.shape[0], x2.shape[1], float(eps), 0,
         _stream())
    return out.view(x.shape)


_freq_cache: Dict[Tuple, torch.Tensor] = {}


def rope_freq_table(rotary_dim: int, rope_scale: float, rope_theta: float, interleave: bool,
                    low_freq_factor=None, high_freq_factor=None, old_context_len=None,
                    device=None) -> torch.Tensor:
    """Per-element frequency table (fp32, [rotary_dim]) with FlashInfer's formula
    (pos_enc.cuh:594-602, 1399-1400, 1538-1539), computed once on the device by vb_rope_freqs."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    llama31 = any(v is not None for v in (low_freq_factor, high_freq_factor, old_context_len))
    lo = 1.0 if low_freq_factor is None else float(low_freq_factor)
    hi = 4.0 if high_freq_factor is None else float(high_freq_factor)
    ctx = 8192.0 if old_context_len is None else float(old_context_len)
    key = (int(rotary_dim), float(rope_scale), float(rope_theta), bool(interleave), llama31, lo, hi, ctx, str(device))
    t = _freq_cache.get(key)
    if t is None:
        t = torch.empty(rotary_dim, dtype=torch.float32, device=device)
        call("vb_rope_freqs", t.data_ptr(), int(rotary_dim), int(bool(interleave)), float(rope_scale),
             float(rope_theta), int(llama31), lo, hi, ctx, _stream())
        _freq_cache[key] = t
    return t


def rope(q: torch.Tensor, k: torch.Tensor, pos: torch.Tensor, freq: torch.Tensor, interleave: bool = False,
         inplace: bool = False):
    """q [T, Hq, D], k [T, Hkv, D] bf16; pos [T] int32."""
    _need_cuda(q, k, pos, freq)
    assert q.dtype == BF16 and k.dtype == BF16 and pos.dtype == torch.int32
    q, k = q.contiguous(), k.contiguous()
    qo = q if inplace else torch.empty_like(q)
    ko = k if inplace else torch.empty_like(k)
    call("vb_rope", qo.data_ptr(), ko.data_ptr(), q.data_ptr(), k.data_ptr(), pos.data_ptr(), freq.data_ptr(),
         q.shape[0], q.shape[1], k.shape[1], q.shape[2], freq.numel(), int(bool(interleave)), _stream())
    return qo, ko


def attn_chunk_tokens(page_size: int, n_kv: int = 8) -> int:
    """Tokens per attention tile (all kv heads ride in one tile): vb_attn_tile_tokens."""
    t = _lib.load().vb_attn_tile_tokens(int(page_size), int(n_kv))
    if t <= 0:
        raise VoxB200Error(f"page_size {page_size} / {n_kv} kv heads unsupported: page_size must be a multiple of 16, "
                           "kv heads <= 8")
    return t


class RowPlan:
    """Device-side per-row metadata produced by vb_plan_rows (see include/vb_api.h); also remembers the page
    table (kv_indices) the plan was made from, which the attention kernel reads."""

    def __init__(self, max_rows: int, device, max_chunks: Optional[int] = None):
        self.max_rows = max_rows
        self.buf = torch.zeros(7 * max_rows + 8, dtype=torch.int32, device=device)
        m = max_rows
        self.row_req, self.row_kvlen = self.buf[0:m], self.buf[m:2 * m]
        self.row_page, self.row_slot = self.buf[2 * m:3 * m], self.buf[3 * m:4 * m]
        self.row_pagebase = self.buf[4 * m:5 * m]
        self.row_old = self.buf[5 * m:6 * m]
        self.row_chunk_start = self.buf[6 * m:7 * m + 1]
        self.kv_indices: Optional[torch.Tensor] = None
        self.qo_indptr: Optional[torch.Tensor] = None      # None: a decode plan (one row per request)
        self.kv_indptr: Optional[torch.Tensor] = None
        self.n_rows = 0
        self.n_req = 0


def plan_rows(plan: RowPlan, qo_indptr: Optional[torch.Tensor], kv_indptr: torch.Tensor, kv_indices: torch.Tensor,
              last_page_len: Optional[torch.Tensor], n_req: int, n_rows: int, page_size: int, chunk_tokens: int,
              kv_len: Optional[torch.Tensor] = None) -> RowPlan:
    _need_cuda(kv_indptr, kv_indices, last_page_len, kv_len)
    assert n_rows <= plan.max_rows
    call("vb_plan_rows", _p(qo_indptr), kv_indptr.data_ptr(), kv_indices.data_ptr(), _p(last_page_len), _p(kv_len),
         n_req, n_rows, page_size, chunk_tokens, plan.row_req.data_ptr(), plan.row_kvlen.data_ptr(),
         plan.row_page.data_ptr(), plan.row_slot.data_ptr(), plan.row_chunk_start.data_ptr(),
         plan.row_pagebase.data_ptr(), plan.row_old.data_ptr(), _stream())
    plan.n_rows = n_rows
    plan.n_req = n_req
    plan.kv_indices = kv_indices
    plan.qo_indptr, plan.kv_indptr = qo_indptr, kv_indptr
    return plan


def kv_append(layer_kv: torch.Tensor, k: torch.Tensor, v: torch.Tensor, plan: RowPlan, n_rows: Optional[int] = None):
    _need_cuda(layer_kv, k, v)
    T = k.shape[0] if n_rows is None else n_rows
    k, v = k.contiguous(), v.contiguous()
    call("vb_kv_append", layer_kv.data_ptr(), k.data_ptr(), v.data_ptr(), plan.row_page.data_ptr(),
         plan.row_slot.data_ptr(), T, layer_kv.shape[-3], layer_kv.shape[-2], layer_kv.shape[-1], _stream())


def copy_pages(kv_cache: torch.Tensor, page_ids: torch.Tensor, staging: Optional[torch.Tensor] = None,
               to_cache: bool = False) -> torch.Tensor:
    """Whole pages of every layer between the paged cache [L, pages, 2, P, Hkv, D] and a contiguous staging buffer
    [L, n, 2, P, Hkv, D] (vb_copy_pages): ``to_cache=False`` gathers (the sender of a KV hand-off), ``True`` scatters
    (the receiver).  page_ids: int32 on the device."""
    _need_cuda(kv_cache, page_ids)
    assert kv_cache.is_contiguous() and kv_cache.dim() == 6 and page_ids.dtype == torch.int32
    L, P = kv_cache.shape[0], kv_cache.shape[1]
    n = page_ids.numel()
    if staging is None:
        assert not to_cache
        staging = torch.empty((L, n) + tuple(kv_cache.shape[2:]), dtype=kv_cache.dtype, device=kv_cache.device)
    _need_cuda(staging)
    assert staging.is_contiguous() and staging.dtype == kv_cache.dtype and staging.numel() == L * n * kv_cache[0, 0].numel()
    page_bytes = kv_cache[0, 0].numel() * kv_cache.element_size()
    call("vb_copy_pages", kv_cache.data_ptr(), staging.data_ptr(), page_ids.data_ptr(), n, L, P, page_bytes,
         int(bool(to_cache)), _stream())
    return staging


def attn_grid_ctas() -> int:
    """Persistent grid of the attention kernel: one CTA per SM (a 3-stage ring of 64 KiB tiles each)."""
    return device_info()[0]


class AttnWorkspace:
    """Split-KV partials + arrival counters, sized for (max_rows, grid CTAs)."""

    def __init__(self, max_rows: int, n_q: int, n_kv: int, head_dim: int, device, grid_ctas: Optional[int] = None):
        self.grid = attn_grid_ctas() if grid_ctas is None else int(grid_ctas)
        self.max_rows = max_rows
        n = _lib.load().vb_paged_attn_workspace_bytes(max_rows, self.grid, n_q, n_kv, head_dim)
        self.buf = torch.zeros(n, dtype=torch.uint8, device=device)


def paged_attn_workspace(max_rows: int, max_chunks, n_q: int, n_kv: int, head_dim: int, device,
                         grid_ctas: Optional[int] = None) -> AttnWorkspace:
    return AttnWorkspace(max_rows, n_q, n_kv, head_dim, device, grid_ctas)


def prefill_attn_tile_rows(n_q: int, n_kv: int) -> int:
    """Rows of one request that share a K/V tile in the tiled prefill kernel (vb_prefill_attn_tile_rows)."""
    return _lib.load().vb_prefill_attn_tile_rows(int(n_q), int(n_kv))


_PREFILL_TILES = os.environ.get("VB_PREFILL_TILES", "auto")      # "0": never, "1": every prefill plan, "auto": by shape
# which tiled prefill kernel: "auto" = tcgen05 for long prompts and for big batches of prompts (measured per layer: 4 x 600
# rows 61 vs 75 us, 8 x 133 rows 21 vs 29 us; but one 133-row prompt 15 vs 12 us, 16 x 50 rows 13 vs 10 us -- the tcgen05
# version pays a V transpose and serial MMA -> softmax -> MMA phases per K/V tile), "1" = always tcgen05, "0" = always mma.sync
_PREFILL_TC = os.environ.get("VB_PREFILL_TC", "auto")


def use_prefill_tc(plan: "RowPlan", n_rows: int) -> bool:
    if _PREFILL_TC in ("0", "1"):
        return _PREFILL_TC == "1"
    n_req = max(1, plan.n_req)
    return n_rows >= 512 * n_req or (n_rows >= 1024 and n_rows >= 128 * n_req)


def use_prefill_tiles(plan: RowPlan, n_rows: int, head_dim: int, page_size: int) -> bool:
    """Tiled prefill kernel or one KV stream per row?  A host-side decision on shapes only (stable under CUDA-graph
    capture): the step must be a prefill plan whose rows are mostly prompt rows -- a joining prompt riding with a batch
    of decodes (133 + 31 rows) qualifies, the 2-row depth-decoder prefill of a multi-codebook frame does not."""
    if plan.qo_indptr is None or _PREFILL_TILES == "0" or head_dim not in (64, 128) or page_size % 16 != 0:
        return False
    if _PREFILL_TILES == "1":
        return True
    return n_rows >= 4 * plan.n_req or n_rows - plan.n_req >= 96


def paged_attn(q: torch.Tensor, kv_cache, slab_base: int, plan: RowPlan, n_rows: int, n_kv: int,
               page_size: int, chunk_tokens: int, workspace: AttnWorkspace, sm_scale: Optional[float] = None,
               out: Optional[torch.Tensor] = None, grid_ctas: Optional[int] = None,
               prefill_tiles: Optional[bool] = None) -> torch.Tensor:
    """q [R, Hq, D] bf16 -> [R, Hq, D].  `kv_cache`: the whole cache tensor [L, pages, 2, P, Hkv, D] (or one layer
    [pages, 2, P, Hkv, D]); slab_base = layer * pages.  `plan` must come from plan_rows with the same
    chunk_tokens (= attn_chunk_tokens(page_size, n_kv)).  Prefill-shaped plans (use_prefill_tiles, or
    ``prefill_tiles=True``) run on the tiled tensor-core kernel (vb_paged_prefill_attn), everything else on the
    one-stream-per-row kernel (vb_paged_attn).  ``prefill_tiles="tc"`` selects the tcgen05 / TMEM variant of the tiled
    kernel (vb_paged_prefill_attn_tc)."""
    if isinstance(kv_cache, TensorMap):      # older call sites pass tensor_map_kv(...): use the tensor behind it
        kv_cache = kv_cache.owner
    _need_cuda(q, kv_cache)
    assert kv_cache.dtype == BF16 and kv_cache.is_contiguous()
    assert q.dtype == BF16 and q.is_contiguous() and plan.kv_indices is not None
    n_q, d = q.shape[1], q.shape[2]
    xt = 0
    if isinstance(out, TiledAct):          # [rows, n_q * d] in the tiled layout: the O projection's activation
        assert out.T == n_rows and out.K == n_q * d
        out_ptr, xt = out.data.data_ptr(), out.t_tile
    else:
        out = torch.empty_like(q) if out is None else out
        out_ptr = out.data_ptr()
    sc = 1.0 / math.sqrt(d) if sm_scale is None else float(sm_scale)
    tiles = use_prefill_tiles(plan, n_rows, d, page_size) if prefill_tiles is None else bool(prefill_tiles)
    if tiles:
        assert plan.qo_indptr is not None, "the tiled prefill kernel needs a prefill plan (qo_indptr)"
        fn = "vb_paged_prefill_attn_tc" if (prefill_tiles == "tc" or (prefill_tiles is None and use_prefill_tc(plan, n_rows))) \
            else "vb_paged_prefill_attn"
        call(fn, out_ptr, q.data_ptr(), kv_cache.data_ptr(), int(slab_base),
             plan.qo_indptr.data_ptr(), plan.kv_indptr.data_ptr(), plan.kv_indices.data_ptr(),
             plan.row_kvlen.data_ptr(), plan.n_req, n_rows, n_q, n_kv, d, page_size, sc, xt, _stream())
        return out
    grid = workspace.grid if grid_ctas is None else min(int(grid_ctas), workspace.grid)
    call("vb_paged_attn", out_ptr, q.data_ptr(), kv_cache.data_ptr(), int(slab_base), plan.row_kvlen.data_ptr(),
         plan.row_chunk_start.data_ptr(), plan.row_pagebase.data_ptr(), plan.row_old.data_ptr(),
         plan.kv_indices.data_ptr(), n_rows, n_q, n_kv,
         d, page_size, chunk_tokens, sc, workspace.buf.data_ptr(), workspace.buf.numel(), grid, workspace.grid, xt,
         _stream())
    return out


# ----------------------------------------------------------------------------------------------
# projections
# ----------------------------------------------------------------------------------------------
def gemm_t_tile(T: int) -> int:
    return _lib.load().vb_gemm_t_tile(int(T))


def choose_split_k(N: int, K: int, T: int, sms: Optional[int] = None) -> int:
    """Spread a skinny (weight-streaming) projection over all SMs: CTAs = ceil(N/128) * split_k."""
    sms = device_info()[0] if sms is None else sms
    tiles = (N + 127) // 128 * ((T + 255) // 256)
    num_kb = (K + 63) // 64
    if tiles >= sms:
        return 1
    # measured on Orpheus-3B at batch 32 (o: 24 tiles x 48 k-blocks, down: 24 x 128): about a dozen k-blocks per CTA
    # and up to two co-resident CTAs per SM beat "one CTA per SM" (down 6 -> 8, o 6 -> 4: forward 2.641 -> 2.609 ms);
    # an even split keeps every CTA's ring the same length
    limit = max(1, min(num_kb // 12, 2 * sms // tiles))
    even = max(s for s in range(1, limit + 1) if num_kb % s == 0)
    return even if 2 * even > limit else limit


class PackedWeight:
    """A projection weight [N, K] re-tiled for the GEMM kernel (vb_pack_weight_tiles): [n_tile][k_block][tile_rows][64]
    bf16, pre-swizzled, every (tile, k-block) operand one contiguous run in HBM."""
    __slots__ = ("data", "N", "K", "tile_rows")

    def __init__(self, data: torch.Tensor, N: int, K: int, tile_rows: int):
        self.data, self.N, self.K, self.tile_rows = data, N, K, tile_rows

    @property
    def shape(self):
        return (self.N, self.K)

    def logical_bytes(self) -> int:
        return 2 * self.N * self.K


def weight_tiles_bytes(N: int, K: int, tile_rows: int) -> int:
    n = _lib.load().vb_weight_tiles_bytes(int(N), int(K), int(tile_rows))
    assert n > 0, (N, K, tile_rows)
    return n


def pack_weight(w: torch.Tensor, tile_rows: int = 128, out: Optional[torch.Tensor] = None) -> PackedWeight:
    """out: a uint8 slice of exactly weight_tiles_bytes(N, K, tile_rows) bytes (e.g. of the engine's weight arena)."""
    _need_cuda(w)
    assert w.dtype == BF16 and w.dim() == 2 and w.stride(1) == 1
    N, K = w.shape
    n = weight_tiles_bytes(N, K, tile_rows)
    if out is None:
        dst = torch.empty(n, dtype=torch.uint8, device=w.device)
    else:
        assert out.dtype == torch.uint8 and out.numel() == n and out.is_contiguous() and out.data_ptr() % 1024 == 0
        dst = out
    call("vb_pack_weight_tiles", dst.data_ptr(), w.data_ptr(), N, K, w.stride(0), tile_rows, _stream())
    return PackedWeight(dst, N, K, tile_rows)


_pack_cache: Dict[Tuple, Tuple[PackedWeight, torch.Tensor]] = {}


def _packed(w, tile_rows: int) -> PackedWeight:
    """PackedWeight as is; a plain [N, K] tensor is packed on first use (convenience for tests / one-off calls --
    the engine packs at load time and drops the row-major copy)."""
    if isinstance(w, PackedWeight):
        assert w.tile_rows == tile_rows, f"weight packed for tile_rows {w.tile_rows}, kernel wants {tile_rows}"
        return w
    key = (w.data_ptr(), tuple(w.shape), w.stride(0), tile_rows, w._version)
    hit = _pack_cache.get(key)
    if hit is None:
        if len(_pack_cache) > 16:
            _pack_cache.clear()
        # the entry keeps the SOURCE tensor alive: a freed tensor's address can be handed to a different weight of the
        # same shape, which would otherwise hit this key and get the old packed copy
        hit = _pack_cache[key] = (pack_weight(w, tile_rows), w)
    return hit[0]


def gemm(x, w, mode: int = 0, split_k: int = 1, out=None, tile_rows: int = 128, n_out: Optional[int] = None,
         bias: Optional[torch.Tensor] = None):
    """x [T, K] bf16, w [N, K] bf16 (nn.Linear.weight layout, or a PackedWeight).  bias (bf16 [N], mode 0 only) is added
    in fp32 before the single bf16 rounding.
    mode 0 -> bf16 [T, N]; mode 1 -> fp32 partials [split_k, T, N]; mode 2 -> bf16 [T, n_out] = silu(gate)*up
    with w rows packed per tile_rows-row tile (see interleave_gate_up; n_out defaults to N/2)."""
    pw = _packed(w, tile_rows)
    if isinstance(x, TiledAct):
        T, K = x.T, x.K
        assert K == pw.K and x.t_tile == gemm_t_tile(T)
        x_map_ptr, x_tiles_ptr, dev = None, x.data.data_ptr(), x.data.device
    else:
        _need_cuda(x)
        assert x.dtype == BF16 and x.dim() == 2 and x.shape[1] == pw.K
        T, K = x.shape
        x_map_ptr, x_tiles_ptr, dev = tensor_map_2d(x, gemm_t_tile(T)).ptr, None, x.device
    N = pw.N
    n_out = (N // 2 if mode == 2 else N) if n_out is None else n_out
    if isinstance(out, TiledAct):
        assert mode == 2 and out.T == T and out.K == n_out
        call("vb_gemm_bf16", out.data.data_ptr(), pw.data.data_ptr(), x_map_ptr, x_tiles_ptr, T, N, K, n_out, mode, split_k,
             tile_rows, n_out, 1, None, _stream())
        return out
    if out is None:
        if mode == 1:
            out = torch.empty(split_k, T, n_out, dtype=torch.float32, device=dev)
        else:
            out = torch.empty(T, n_out, dtype=BF16, device=dev)
    if bias is not None:
        _need_cuda(bias)
        assert mode == 0 and bias.dtype == BF16 and bias.numel() == N and bias.is_contiguous()
    call("vb_gemm_bf16", out.data_ptr(), pw.data.data_ptr(), x_map_ptr, x_tiles_ptr, T, N, K, n_out, mode, split_k,
         tile_rows, n_out, 0, _p(bias), _stream())
    return out


def norm_lmhead(hidden: torch.Tensor, norm_weight: torch.Tensor, eps: float, w, bias: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """logits = lm_head(rmsnorm(hidden) * norm_weight) (+ bias) as one C-ABI operator (vb_norm_lmhead; orpheus.py:193-197,
    219-221).  hidden [T, K] bf16, w [N, K] (or a PackedWeight) -> bf16 [T, N]."""
    pw = _packed(w, 128)
    _need_cuda(hidden, norm_weight)
    assert hidden.dtype == BF16 and hidden.dim() == 2 and hidden.is_contiguous() and hidden.shape[1] == pw.K
    T, K = hidden.shape
    n = _lib.load().vb_norm_lmhead_workspace_bytes(T, K)
    if workspace is None:
        workspace = torch.empty(n, dtype=torch.uint8, device=hidden.device)
    assert workspace.numel() * workspace.element_size() >= n
    out = torch.empty(T, pw.N, dtype=BF16, device=hidden.device) if out is None else out
    call("vb_norm_lmhead", out.data_ptr(), hidden.data_ptr(), norm_weight.data_ptr(), float(eps), pw.data.data_ptr(),
         _p(bias), T, pw.N, K, out.stride(0), pw.tile_rows, workspace.data_ptr(),
         workspace.numel() * workspace.element_size(), _stream())
    return out


def gate_up_tile_half(I: int, sms: Optional[int] = None) -> int:
    """Gate rows per tile (h): a multiple of 16 (each epilogue warp pairs 16 gate with 16 up rows), at most 64, chosen for
    the fullest waves: up to one CTA per SM runs with the whole-SM ring, more tiles than SMs run two per SM (the kernel
    then takes the 104 KB ring).  Orpheus / CSM (8192): 64 -> 128 CTAs; GLM-4-Voice (13696): 48 -> 286 CTAs on 296 slots
    (64 would leave 214 on 296: measured 6.28 -> 5.75 ms per step); CosyVoice2 (4864): 48; Qwen3-TTS (6144): 48."""
    sms = device_info()[0] if sms is None else sms
    best, best_eff = 64, -1.0
    for h in (64, 48, 32, 16):
        tiles = (I + h - 1) // h
        slots = sms if tiles <= sms else 2 * sms
        eff = tiles / float((tiles + slots - 1) // slots * slots)
        if eff > best_eff + 1e-9:
            best, best_eff = h, eff
    return best


def interleave_gate_up(gate_w: torch.Tensor, up_w: torch.Tensor, h: int = 64) -> torch.Tensor:
    """[I, K] x 2 -> [2 * ceil(I/h) * h, K].  A tile holds h gate rows and the h matching up rows, grouped per
    epilogue warp: [16 gate rows][their 16 up rows] x (h / 16), so that gate and up of one output sit 16 lanes apart in
    the same warp's accumulator quarter.  I is zero-padded to a multiple of h (the padded outputs are never stored)."""
    I, K = gate_w.shape
    assert up_w.shape == gate_w.shape and h % 16 == 0 and 16 <= h <= 64
    tiles = (I + h - 1) // h
    if tiles * h != I:
        pad = torch.zeros(tiles * h - I, K, dtype=gate_w.dtype, device=gate_w.device)
        gate_w, up_w = torch.cat((gate_w, pad), 0), torch.cat((up_w, pad), 0)
    g = gate_w.view(tiles * (h // 16), 16, K)
    u = up_w.view(tiles * (h // 16), 16, K)
    return torch.cat((g, u), dim=1).reshape(2 * tiles * h, K).contiguous()


def fused_t_tile(T: int) -> int:
    return 16 if T <= 16 else (32 if T <= 32 else 64)


def proj_split_k(n_tiles: int, K: int, sms: Optional[int] = None) -> int:
    """split-K of a fused (cluster-reduced) projection: about one CTA per SM, at most 8 CTAs per tile, at least
    4 K-blocks per CTA."""
    sms = device_info()[0] if sms is None else sms
    num_kb = (K + 63) // 64
    return max(1, min(8, sms // max(1, n_tiles), max(1, num_kb // 4)))


def _act_ptrs(x, t_tile: int):
    """(x_map ptr or None, x_tiles ptr or None, T, K, device) of a row-major tensor or a TiledAct"""
    if isinstance(x, TiledAct):
        assert x.t_tile == t_tile
        return None, x.data.data_ptr(), x.T, x.K, x.data.device
    _need_cuda(x)
    assert x.dtype == BF16 and x.dim() == 2
    return tensor_map_2d(x, t_tile).ptr, None, x.shape[0], x.shape[1], x.device


def proj_residual(x, w, residual: Optional[torch.Tensor], split_k: int,
                  hidden_out: Optional[torch.Tensor] = None, ssq_out: Optional[torch.Tensor] = None,
                  tile_rows: int = 128, hidden_tiles_out: Optional["TiledAct"] = None):
    """hidden_out [T, N] = bf16(residual + bf16(x w^T)); ssq_out fp32 [ceil(N/tile_rows), T] per-tile sums of squares
    of the new rows; hidden_tiles_out: the same rows once more in the tiled layout (the next norm-fused projection's
    input).  x: [T, K] tensor or TiledAct.  T <= 64."""
    pw = _packed(w, tile_rows)
    T = x.T if isinstance(x, TiledAct) else x.shape[0]
    x_map, x_tiles, T, K, dev = _act_ptrs(x, fused_t_tile(T))
    N = pw.N
    tiles = (N + tile_rows - 1) // tile_rows
    hidden_out = torch.empty(T, N, dtype=BF16, device=dev) if hidden_out is None else hidden_out
    ssq_out = torch.empty(tiles, T, dtype=torch.float32, device=dev) if ssq_out is None else ssq_out
    assert ssq_out.numel() >= tiles * T and K == pw.K
    if hidden_tiles_out is not None:
        assert hidden_tiles_out.T == T and hidden_tiles_out.K == N
    call("vb_proj_residual", hidden_out.data_ptr(), hidden_tiles_out.data.data_ptr() if hidden_tiles_out is not None else None,
         ssq_out.data_ptr(), pw.data.data_ptr(), x_map, x_tiles, _p(residual), T, N, K, split_k, tile_rows, _stream())
    return hidden_out, ssq_out


def proj_norm_gateup_silu(hidden, ssq: torch.Tensor, n_parts: int, norm_w: torch.Tensor, eps: float,
                          w_packed, h: int, n_out: int, out=None):
    """act [T, n_out] = silu(gate(xn)) * up(xn) with xn = rmsnorm(hidden) * norm_w formed inside the kernel.
    hidden: [T, K] tensor or TiledAct; out: tensor, TiledAct or None."""
    _need_cuda(ssq, norm_w)
    pw = _packed(w_packed, 2 * h)
    T = hidden.T if isinstance(hidden, TiledAct) else hidden.shape[0]
    x_map, x_tiles, T, K, dev = _act_ptrs(hidden, fused_t_tile(T))
    tiled_out = isinstance(out, TiledAct)
    if out is None:
        out = torch.empty(T, n_out, dtype=BF16, device=dev)
    call("vb_proj_norm_gateup_silu", out.data.data_ptr() if tiled_out else out.data_ptr(), pw.data.data_ptr(), x_map, x_tiles,
         ssq.data_ptr(), n_parts, norm_w.data_ptr(), float(eps), T, pw.N, K, 2 * h, n_out, 1 if tiled_out else 0, _stream())
    return out


def proj_norm_qkv_rope_append(hidden, ssq: torch.Tensor, n_parts: int, norm_w: torch.Tensor, eps: float,
                              w_qkv, layer_kv: torch.Tensor, rope_cs: torch.Tensor, plan: "RowPlan",
                              n_q: int, n_kv: int, head_dim: int, split_k: int,
                              q_out: Optional[torch.Tensor] = None):
    """q [T, n_q, D] (rotated) and the rotated k / v of every row written to its page slot; see vb_api.h.
    hidden: [T, K] tensor or TiledAct."""
    _need_cuda(ssq, norm_w, layer_kv, rope_cs)
    pw = _packed(w_qkv, head_dim)
    T = hidden.T if isinstance(hidden, TiledAct) else hidden.shape[0]
    x_map, x_tiles, T, K, dev = _act_ptrs(hidden, fused_t_tile(T))
    assert pw.N == (n_q + 2 * n_kv) * head_dim
    q_out = torch.empty(T, n_q, head_dim, dtype=BF16, device=dev) if q_out is None else q_out
    page_size = layer_kv.shape[-3]
    call("vb_proj_norm_qkv_rope_append", q_out.data_ptr(), layer_kv.data_ptr(), pw.data.data_ptr(), x_map, x_tiles,
         ssq.data_ptr(), n_parts, norm_w.data_ptr(), float(eps), rope_cs.data_ptr(), plan.row_page.data_ptr(),
         plan.row_slot.data_ptr(), T, K, n_q, n_kv, head_dim, page_size, split_k, _stream())
    return q_out


def rope_table(pos: torch.Tensor, freq: torch.Tensor, head_dim: int, out: Optional[torch.Tensor] = None):
    """cos | sin of pos[t] * freq[e]: fp32 [T, 2, D], shared by every layer of the step."""
    _need_cuda(pos, freq)
    T = pos.numel()
    out = torch.empty(T, 2, head_dim, dtype=torch.float32, device=pos.device) if out is None else out
    call("vb_rope_table", out.data_ptr(), pos.data_ptr(), freq.data_ptr(), T, head_dim, _stream())
    return out


def row_ssq(x: torch.Tensor, out: Optional[torch.Tensor] = None):
    _need_cuda(x)
    T, dim = x.shape
    out = torch.empty(T, dtype=torch.float32, device=x.device) if out is None else out
    call("vb_row_ssq", out.data_ptr(), x.data_ptr(), T, dim, _stream())
    return out


def reduce_residual_rmsnorm(partials: torch.Tensor, residual: Optional[torch.Tensor],
                            norm_weight: Optional[torch.Tensor], eps: float,
                            hidden_out: Optional[torch.Tensor] = None, normed_out: Optional[torch.Tensor] = None,
                            want_hidden: bool = True):
    """partials fp32 [S, T, N] -> (hidden bf16 [T, N], normed bf16 [T, N] or None)."""
    _need_cuda(partials)
    S, T, N = partials.shape
    if want_hidden and hidden_out is None:
        hidden_out = torch.empty(T, N, dtype=BF16, device=partials.device)
    xt = 0
    if isinstance(normed_out, TiledAct):
        assert normed_out.T == T and normed_out.K == N and norm_weight is not None
        n_ptr, xt = normed_out.data.data_ptr(), normed_out.t_tile
    else:
        if norm_weight is not None and normed_out is None:
            normed_out = torch.empty(T, N, dtype=BF16, device=partials.device)
        n_ptr = _p(normed_out)
    call("vb_reduce_residual_rmsnorm", _p(hidden_out), n_ptr, partials.data_ptr(), S, _p(residual),
         _p(norm_weight), T, N, float(eps), xt, _stream())
    return hidden_out, normed_out


def qkv_rope_append(partials: torch.Tensor, layer_kv: torch.Tensor, pos: torch.Tensor, freq: torch.Tensor,
                    plan: RowPlan, n_q: int, n_kv: int, head_dim: int, interleave: bool = False,
                    q_out: Optional[torch.Tensor] = None, q_norm: Optional[torch.Tensor] = None,
                    k_norm: Optional[torch.Tensor] = None, norm_eps: float = 1e-6,
                    qkv_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q_norm / k_norm (bf16 [head_dim]): per-head RMSNorm of q and k before the rotation (Qwen3).  qkv_bias (bf16
    [(n_q + 2 n_kv) head_dim]): bias of the q | k | v projections (CosyVoice2, GLM-4-Voice).  freq may hold fewer than
    head_dim entries: only the first freq.numel() elements of every head are rotated (GLM: half of the head)."""
    _need_cuda(partials, layer_kv, pos, freq, q_norm, k_norm, qkv_bias)
    assert qkv_bias is None or (qkv_bias.dtype == BF16 and qkv_bias.numel() == (n_q + 2 * n_kv) * head_dim)
    S, T, W = partials.shape
    assert W == (n_q + 2 * n_kv) * head_dim
    if q_out is None:
        q_out = torch.empty(T, n_q, head_dim, dtype=BF16, device=partials.device)
    call("vb_qkv_rope_append", q_out.data_ptr(), layer_kv.data_ptr(), partials.data_ptr(), S, pos.data_ptr(),
         freq.data_ptr(), plan.row_page.data_ptr(), plan.row_slot.data_ptr(), T, n_q, n_kv, head_dim,
         layer_kv.shape[-3], freq.numel(), int(bool(interleave)), _p(q_norm), _p(k_norm), float(norm_eps), _p(qkv_bias),
         _stream())
    return q_out


def embedding(table: torch.Tensor, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(table, ids)
    assert ids.dtype == torch.int32 and table.dtype == BF16
    T = ids.numel()
    out = torch.empty(T, table.shape[1], dtype=BF16, device=table.device) if out is None else out
    call("vb_embedding", out.data_ptr(), table.data_ptr(), ids.data_ptr(), T, table.shape[1], table.shape[0], _stream())
    return out


def multi_embed_sum(out: torch.Tensor, ids: torch.Tensor, table_a: Optional[torch.Tensor], col_offset: int = 0,
                    col0: int = 0, n_cols_a: Optional[int] = None, table_b: Optional[torch.Tensor] = None,
                    mask: Optional[torch.Tensor] = None, round_each: bool = False) -> torch.Tensor:
    """out [T, dim] bf16 = masked sum over the C columns of ids [T, C] (int64, any strides) of embedding rows:
    columns < n_cols_a from table_a at row ids + (col0 + c) * col_offset, the rest from table_b.  mask [T, C] bool/uint8
    contiguous or None.  (vb_multi_embed_sum; csm.py:637-663)"""
    _need_cuda(out, ids, table_a, table_b, mask)
    assert ids.dtype == torch.int64 and ids.dim() == 2 and out.dtype == BF16 and out.dim() == 2 and out.stride(1) == 1
    T, C = ids.shape
    n_cols_a = C if n_cols_a is None else n_cols_a
    m8 = None
    if mask is not None:
        assert mask.shape == (T, C) and mask.is_contiguous()
        m8 = mask.view(torch.uint8) if mask.dtype == torch.bool else mask
    dim = out.shape[1]
    for tb in (table_a, table_b):
        assert tb is None or (tb.dtype == BF16 and tb.is_contiguous() and tb.shape[1] == dim)
    call("vb_multi_embed_sum", out.data_ptr(), out.stride(0), ids.data_ptr(), ids.stride(0), ids.stride(1), _p(m8),
         _p(table_a), table_a.shape[0] if table_a is not None else 0, int(col_offset), int(col0), int(n_cols_a),
         _p(table_b), table_b.shape[0] if table_b is not None else 0, T, C, dim, int(bool(round_each)), _stream())
    return out


def talker_embed(out: torch.Tensor, text: torch.Tensor, codec: torch.Tensor, cb0: torch.Tensor,
                 needs_codec: Optional[torch.Tensor] = None, features: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Qwen3-TTS talker input rows (vb_talker_embed).  text: [T, H] or [H] (broadcast); cb0 int64 [T] (any stride);
    needs_codec bool [T] or None (all); features [T, H] or None."""
    _need_cuda(out, text, codec, cb0, needs_codec, features)
    T, H = out.shape
    assert cb0.dtype == torch.int64 and cb0.dim() == 1 and cb0.numel() == T and codec.is_contiguous()
    ld_text = 0 if text.dim() == 1 else text.stride(0)
    nc = None
    if needs_codec is not None:
        assert needs_codec.numel() == T and needs_codec.is_contiguous()
        nc = needs_codec.view(torch.uint8) if needs_codec.dtype == torch.bool else needs_codec
    call("vb_talker_embed", out.data_ptr(), out.stride(0), text.data_ptr(), ld_text, codec.data_ptr(), codec.shape[0],
         cb0.data_ptr(), cb0.stride(0), _p(nc), _p(features), features.stride(0) if features is not None else 0, T, H,
         _stream())
    return out


def interleave_rows(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n, d] x 2 -> [2n, d] with rows a0, b0, a1, b1, ..."""
    _need_cuda(a, b)
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous() and a.dim() == 2
    n, d = a.shape
    out = torch.empty(2 * n, d, dtype=a.dtype, device=a.device) if out is None else out
    call("vb_interleave_rows", out.data_ptr(), a.data_ptr(), b.data_ptr(), n, d * a.element_size(), _stream())
    return out


def transpose_i64(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src int64 [C, B] (row stride arbitrary) -> [B, C]"""
    _need_cuda(src)
    assert src.dtype == torch.int64 and src.dim() == 2 and src.stride(1) == 1
    C, B = src.shape
    out = torch.empty(B, C, dtype=torch.int64, device=src.device) if out is None else out
    assert out.stride(1) == 1
    call("vb_transpose_i64", out.data_ptr(), src.data_ptr(), B, C, out.stride(0), src.stride(0), _stream())
    return out


def gather_rows(src: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None,
                idx_offset: int = 0) -> torch.Tensor:
    _need_cuda(src, idx)
    assert idx.dtype == torch.int32 and src.dim() == 2 and src.is_contiguous()
    out = torch.empty(idx.numel(), src.shape[1], dtype=src.dtype, device=src.device) if out is None else out
    call("vb_gather_rows", out.data_ptr(), src.data_ptr(), idx.data_ptr(), idx.numel(),
         src.shape[1] * src.element_size(), int(idx_offset), _stream())
    return out


# ----------------------------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------------------------
STRATEGY = {"greedy": 0, "top_k": 1, "top_p": 2, "top_k_top_p": 3, "min_p": 4}
_sample_ws: Dict[Tuple, torch.Tensor] = {}


def sample_workspace(rows: int, vocab: int, device) -> torch.Tensor:
    """One workspace per device, sized for at least 64 rows up front so that later (CUDA-graph captured) calls
    never allocate."""
    key = str(device)
    ws = _sample_ws.get(key)
    n = _lib.load().vb_sample_workspace_bytes(max(rows, 64), vocab)
    if ws is None or ws.numel() < n:
        ws = torch.empty(n, dtype=torch.uint8, device=device)
        _sample_ws[key] = ws
    return ws


def _cache_u8(rep_cache: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if rep_cache is None:
        return None
    assert rep_cache.dim() == 4 and rep_cache.is_contiguous()
    return rep_cache.view(torch.uint8) if rep_cache.dtype == torch.bool else rep_cache


def sample(logits: torch.Tensor, strategy: str, rep_cache: Optional[torch.Tensor] = None, penalty: float = 1.0,
           logit_codebooks: int = 1, top_k: int = 0, top_p: float = 1.0, min_p: float = 0.0, temperature: float = 1.0,
           seed: int = 0, offset: int = 0, mask_token: int = -1, out: Optional[torch.Tensor] = None,
           workspace: Optional[torch.Tensor] = None, rng_state: Optional[torch.Tensor] = None,
           cache_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    """logits [rows, V] bf16 (row stride arbitrary) -> ids int64 [rows].
    cache_rows (int32 [batch]): batch row b reads rep_cache[cache_rows[b]] (slot-resident caches)."""
    _need_cuda(logits)
    assert logits.dtype == BF16 and logits.dim() == 2 and logits.stride(1) == 1
    rows, V = logits.shape
    out = torch.empty(rows, dtype=torch.int64, device=logits.device) if out is None else out
    ws = sample_workspace(rows, V, logits.device) if workspace is None else workspace
    if rng_state is not None:
        assert rng_state.dtype == torch.int64 and rng_state.numel() >= 3, "rng_state: int64 {seed, offset, 0}"
    c8 = _cache_u8(rep_cache)
    W, Cc = (c8.shape[1], c8.shape[2]) if c8 is not None else (0, 0)
    call("vb_sample", out.data_ptr(), logits.data_ptr(), rows, V, logits.stride(0), _p(c8), _p(cache_rows), W, Cc,
         logit_codebooks,
         float(penalty), STRATEGY[strategy], int(top_k or 0), float(top_p if top_p is not None else 1.0),
         float(min_p or 0.0), float(temperature), int(seed), int(offset), _p(rng_state), int(mask_token), ws.data_ptr(),
         ws.numel(),
         _stream())
    return out


def apply_repetition_penalty(logits: torch.Tensor, rep_cache: torch.Tensor, penalty: float) -> torch.Tensor:
    """logits [B, C, V] bf16, cache [B, W, Cc, V] bool -> penalised logits (new tensor)."""
    _need_cuda(logits, rep_cache)
    B, Cl, V = logits.shape
    x = logits.contiguous()
    c8 = _cache_u8(rep_cache)
    out = torch.empty_like(x)
    call("vb_apply_repetition_penalty", out.data_ptr(), x.data_ptr(), c8.data_ptr(), c8.shape[1], c8.shape[2], Cl,
         float(penalty), B * Cl, V, _stream())
    return out


def update_repetition_cache(rep_cache: torch.Tensor, ids: torch.Tensor, window: int,
                            cache_rows: Optional[torch.Tensor] = None) -> None:
    """ids int64 [B, C_ids]; with cache_rows the B batch rows mark cache rows cache_rows[b] of a larger
    slot-resident cache."""
    _need_cuda(rep_cache, ids)
    c8 = _cache_u8(rep_cache)
    ids64 = ids if (ids.dtype == torch.int64 and ids.is_contiguous()) else ids.to(torch.int64).contiguous()
    _, W, Cc, V = c8.shape
    B = ids64.shape[0]
    call("vb_update_repetition_cache", c8.data_ptr(), _p(cache_rows), ids64.data_ptr(), B, W, Cc, V, ids64.shape[1],
         int(window), _stream())


# ----------------------------------------------------------------------------------------------
# device-resident decode loop helpers
# ----------------------------------------------------------------------------------------------
def decode_advance(kv_len: torch.Tensor, pos: torch.Tensor, active: Optional[torch.Tensor] = None):
    call("vb_decode_advance", kv_len.data_ptr(), pos.data_ptr(), _p(active), kv_len.numel(), _stream())


def token_feedback(ids: torch.Tensor, slots: Optional[torch.Tensor], next_input: torch.Tensor,
                   history: Optional[torch.Tensor], n_out: torch.Tensor, skip_token: int = -1):
    """ids int64 [B] -> per-slot next input id, history ring [slots, cap], token counter.  An id equal to
    ``skip_token`` (the stop id) is fed back but not recorded: the ring holds audio tokens only."""
    cap = history.shape[1] if history is not None else 1
    call("vb_token_feedback", ids.data_ptr(), _p(slots), next_input.data_ptr(), _p(history), n_out.data_ptr(),
         ids.numel(), cap, int(skip_token), _stream())


def gather_i32(src: torch.Tensor, idx: Optional[torch.Tensor], out: torch.Tensor, n: Optional[int] = None):
    call("vb_gather_i32", out.data_ptr(), src.data_ptr(), _p(idx), out.numel() if n is None else n, _stream())
    return out


def build_input_ids(out: torch.Tensor, host_ids: torch.Tensor, next_input: torch.Tensor, row_slot: torch.Tensor,
                    n: int):
    call("vb_build_input_ids", out.data_ptr(), host_ids.data_ptr(), next_input.data_ptr(), row_slot.data_ptr(), n,
         _stream())
    return out


def latest_window(first: torch.Tensor, n_out: torch.Tensor, slot: Optional[torch.Tensor], n: int, window: int):
    call("vb_latest_window", first.data_ptr(), n_out.data_ptr(), _p(slot), n, window, _stream())
    return first


def gather_windows(history: torch.Tensor, slot: torch.Tensor, first: torch.Tensor, n_valid: Optional[torch.Tensor],
                   window: int, out: Optional[torch.Tensor] = None, n: Optional[int] = None) -> torch.Tensor:
    n = slot.numel() if n is None else n
    out = torch.empty(n, window, dtype=torch.int64, device=history.device) if out is None else out
    call("vb_gather_windows", out.data_ptr(), history.data_ptr(), slot.data_ptr(), first.data_ptr(), _p(n_valid), n,
         history.shape[1], window, _stream())
    return out


# ----------------------------------------------------------------------------------------------
# misc
# ----------------------------------------------------------------------------------------------
def pcm16(audio: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(audio)
    a = audio.contiguous()
    assert a.dtype == torch.float32
    out = torch.empty(a.shape, dtype=torch.int16, device=a.device) if out is None else out
    call("vb_pcm16", out.data_ptr(), a.data_ptr(), a.numel(), _stream())
    return out


def randn(n_or_out, seed: int = 0, offset: int = 0, rng_state: Optional[torch.Tensor] = None,
          device=None) -> torch.Tensor:
    """Standard-normal fp32 draws on the device (vb_randn): ``n_or_out`` = element count or a preallocated fp32
    tensor.  ``rng_state`` (int64 {seed, offset, 0}, device) makes successive calls / graph replays draw fresh values."""
    if isinstance(n_or_out, torch.Tensor):
        out = n_or_out
        _need_cuda(out)
        assert out.dtype == torch.float32 and out.is_contiguous()
    else:
        out = torch.empty(int(n_or_out), dtype=torch.float32,
                          device=device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    if rng_state is not None:
        assert rng_state.dtype == torch.int64 and rng_state.numel() >= 3 and rng_state.is_cuda
    call("vb_randn", out.data_ptr(), out.numel(), int(seed) & ((1 << 64) - 1), int(offset), _p(rng_state), _stream())
    return out


def orpheus_window_codes(ids: torch.Tensor, audio_id_base: int):
    """ids int64 [B, 28] -> codes0 [B,4], codes1 [B,8], codes2 [B,16] int32 (orpheus.py:479-500)."""
    _need_cuda(ids)
    ids = ids.to(torch.int64).contiguous().view(-1, 28)
    B = ids.shape[0]
    c0 = torch.empty(B, 4, dtype=torch.int32, device=ids.device)
    c1 = torch.empty(B, 8, dtype=torch.int32, device=ids.device)
    c2 = torch.empty(B, 16, dtype=torch.int32, device=ids.device)
    call("vb_orpheus_window_codes", c0.data_ptr(), c1.data_ptr(), c2.data_ptr(), ids.data_ptr(), B,
         int(audio_id_base), _stream())
    return c0, c1, c2
