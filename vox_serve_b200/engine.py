"""Llama-shaped LM forward on the sm_100a kernels: the arithmetic of ``OrpheusForCausalLM.forward``
(vox_serve/model/orpheus.py:125-221) as a fixed sequence of C-ABI launches over preallocated buffers, so
the whole step is CUDA-graph capturable (no host sync, no allocation).

Decode-sized steps (<= 64 rows), 5 launches per layer (the reference issues ~18-20):
    [RMSNorm -> QKV projection -> RoPE -> KV append]  ->  paged attention
    -> [O projection -> + residual (+ sums of squares for the next norm)]
    -> [RMSNorm -> gate/up projection -> SiLU * up]  ->  [down projection -> + residual (+ sums of squares)]
Larger steps (prefill), 8 launches per layer: the same projections leaving fp32 split-K partials, with the
reduce / RoPE / append and reduce / residual / RMSNorm tails as separate kernels.
Rounding points follow the reference module graph: every Linear / RMSNorm / RoPE / attention output and
every residual add is rounded to bf16.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._lib import VoxB200Error

BF16 = torch.bfloat16


@dataclass
class LlamaDims:
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    intermediate_size: int
    vocab_size: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_factor: float = 32.0
    low_freq_factor: Optional[float] = 1.0
    high_freq_factor: Optional[float] = 4.0
    old_context_len: Optional[float] = 8192
    qk_norm: bool = False      # per-head RMSNorm of q and k before RoPE (Qwen3: self_attn.q_norm / k_norm)
    qkv_bias: bool = False     # q / k / v projections with bias (CosyVoice2 / Qwen2, GLM-4-Voice)
    head_bias: bool = False    # output head with bias (CosyVoice2's llm_decoder)
    rotary_dim: Optional[int] = None     # elements of every head that are rotated (None = head_dim; GLM: head_dim / 2)
    rope_interleave: bool = False        # rotate (even, odd) pairs instead of the two halves (GLM)

    @classmethod
    def orpheus_3b(cls):
        return cls(3072, 28, 24, 8, 128, 8192, 156940)


def hf_layer_names(i: int, prefix: str = "model.") -> Dict[str, str]:
    p = f"{prefix}layers.{i}."
    return {"ln1": p + "input_layernorm.weight", "ln2": p + "post_attention_layernorm.weight",
            "q": p + "self_attn.q_proj.weight", "k": p + "self_attn.k_proj.weight",
            "v": p + "self_attn.v_proj.weight", "o": p + "self_attn.o_proj.weight",
            "gate": p + "mlp.gate_proj.weight", "up": p + "mlp.up_proj.weight", "down": p + "mlp.down_proj.weight"}


class LlamaWeights:
    """Device-resident weights in the layout the kernels stream: q|k|v rows concatenated, gate/up rows
    packed per tile as h gate rows + the h matching up rows (ops.interleave_gate_up; h is chosen so that the
    projection splits into about one tile per SM); built from an HF-named state dict."""

    def __init__(self, dims: LlamaDims, device="cuda"):
        self.dims, self.device = dims, torch.device(device)
        self.gu_half = ops.gate_up_tile_half(dims.intermediate_size) if self.device.type == "cuda" else 64
        self.gu_half = int(os.environ.get("VB_GU_HALF", self.gu_half))          # (dev: tile-height experiments)
        self.embed = self.norm = self.lm_head = self.head_bias = None
        self.heads: List[object] = []
        self.arena: Optional[torch.Tensor] = None
        self.layers: List[Dict[str, torch.Tensor]] = []

    @classmethod
    def from_state_dict(cls, sd, dims: LlamaDims, device="cuda", prefix: str = "model.",
                        embed_key: Optional[str] = "model.embed_tokens.weight", head_key: str = "lm_head.weight",
                        heads: Optional[List[torch.Tensor]] = None, head_bias_key: Optional[str] = None):
        """``prefix`` / ``embed_key`` / ``head_key``: where the decoder stack, the token embedding (None: the caller
        supplies input embeddings) and the output head live in ``sd`` -- HF Llama names by default; the CSM backbone is
        ``backbone_model.`` + ``lm_head.weight``, its depth decoder ``depth_decoder.model.`` with ``heads`` = one
        [vocab, hidden] matrix per depth position (csm.py:235-255) instead of a single lm_head."""
        self = cls(dims, device)
        dev = self.device

        def put(t):
            return t.to(device=dev, dtype=BF16).contiguous()

        self.embed = put(sd[embed_key]) if embed_key is not None else None
        self.norm = put(sd[prefix + "norm.weight"])
        # every projection of a step lives in ONE arena in the order a decode step streams it (per layer qkv, o,
        # gate/up, down; then lm_head): the L2 prefetcher walks it front to back (ops.weight_prefetch)
        H, I, D = dims.hidden_size, dims.intermediate_size, dims.head_dim
        qkv_rows = (dims.num_attention_heads + 2 * dims.num_key_value_heads) * D
        gu_tiles = (I + self.gu_half - 1) // self.gu_half
        per_layer = [ops.weight_tiles_bytes(qkv_rows, H, D), ops.weight_tiles_bytes(H, dims.num_attention_heads * D, 128),
                     ops.weight_tiles_bytes(2 * gu_tiles * self.gu_half, H, 2 * self.gu_half),
                     ops.weight_tiles_bytes(H, I, 128)]
        head_bytes = ops.weight_tiles_bytes(dims.vocab_size, H, 128)
        n_heads_out = len(heads) if heads is not None else 1
        if dev.type == "cuda":
            raw = torch.empty(dims.num_hidden_layers * sum(per_layer) + n_heads_out * head_bytes + 1024, dtype=torch.uint8,
                              device=dev)
            pad = (-raw.data_ptr()) % 1024           # (small allocations are only 512-byte aligned)
            self.arena = raw[pad:pad + raw.numel() - 1024]
        off = 0
        for i in range(dims.num_hidden_layers):
            n = hf_layer_names(i, prefix)
            slots = []
            for nb in per_layer:
                slots.append(self.arena[off:off + nb] if self.arena is not None else None)
                off += nb
            self.layers.append(self.pack_layer(put(sd[n["ln1"]]), put(sd[n["q"]]), put(sd[n["k"]]), put(sd[n["v"]]),
                                               put(sd[n["o"]]), put(sd[n["ln2"]]), put(sd[n["gate"]]),
                                               put(sd[n["up"]]), put(sd[n["down"]]), self.gu_half, dims.head_dim,
                                               out=slots))
            if dims.qk_norm:
                lp = f"{prefix}layers.{i}.self_attn."
                self.layers[-1]["qn"], self.layers[-1]["kn"] = put(sd[lp + "q_norm.weight"]), put(sd[lp + "k_norm.weight"])
            if dims.qkv_bias:
                lp = f"{prefix}layers.{i}.self_attn."
                self.layers[-1]["qkv_b"] = put(torch.cat([sd[lp + f"{x}_proj.bias"].reshape(-1) for x in "qkv"]))
        if heads is None:
            heads = [sd[head_key] if head_key in sd else sd[embed_key]]
        self.heads = []
        for hw in heads:
            assert tuple(hw.shape) == (dims.vocab_size, H), (tuple(hw.shape), dims.vocab_size, H)
            self.heads.append(ops.pack_weight(put(hw), 128,
                                              out=self.arena[off:off + head_bytes] if self.arena is not None else None))
            off += head_bytes
        self.lm_head = self.heads[0]
        self.head_bias = put(sd[head_bias_key]) if dims.head_bias else None
        if dev.type == "cuda":
            # the row-major originals were just dropped: hand their blocks back NOW, not inside the first CUDA-graph
            # capture on the request path (capture_begin empties the allocator cache: ~1 s for 6.6 GB of blocks)
            torch.cuda.synchronize(dev)
            torch.cuda.empty_cache()
        return self

    @staticmethod
    def pack_layer(ln1, q, k, v, o, ln2, gate, up, down, gu_half: int = 64, head_dim: int = 128,
                   out=None) -> Dict[str, object]:
        """Projection weights leave here re-tiled for the GEMM kernel (ops.pack_weight): one head per QKV tile,
        128-row O / down tiles, gu_half gate + gu_half up rows per gate/up tile.  The row-major copies are dropped.
        out: four uint8 slices (qkv, o, gate/up, down) to pack into, e.g. of the weight arena."""
        out = out or [None] * 4
        return {"ln1": ln1, "ln2": ln2,
                "qkv": ops.pack_weight(torch.cat((q, k, v), 0).contiguous(), head_dim, out=out[0]),
                "o": ops.pack_weight(o, 128, out=out[1]),
                "gu": ops.pack_weight(ops.interleave_gate_up(gate, up, gu_half), 2 * gu_half, out=out[2]),
                "down": ops.pack_weight(down, 128, out=out[3])}

    def nbytes(self) -> int:
        """logical bf16 bytes (N * K * 2 per projection; the packed tiles add only tail padding)"""
        def nb(t):
            return t.logical_bytes() if isinstance(t, ops.PackedWeight) else 2 * t.numel()
        n = (nb(self.embed) if self.embed is not None else 0) + nb(self.norm) + sum(nb(h) for h in self.heads)
        for l in self.layers:
            n += sum(nb(t) for t in l.values())
        return n

    def streamed_bytes_per_step(self) -> int:
        """Weight bytes one decode step must read (everything but the embedding table)."""
        return self.nbytes() - (2 * self.embed.numel() if self.embed is not None else 0)


class LlamaEngine:
    def __init__(self, weights: LlamaWeights, kv_cache: torch.Tensor, page_size: int, max_rows: int,
                 max_seq_len: int = 2304):
        d = weights.dims
        self.w, self.dims, self.kv_cache, self.page_size = weights, d, kv_cache, page_size
        self.device = kv_cache.device
        assert kv_cache.shape[0] == d.num_hidden_layers and kv_cache.shape[2] == 2
        self.pages_per_layer = kv_cache.shape[1]
        self.chunk = ops.attn_chunk_tokens(page_size, d.num_key_value_heads)
        self.kv_map = kv_cache      # the attention kernel reads the cache with linear bulk copies: no tensor map
        self.max_rows = max_rows
        self.sms = ops.device_info()[0]
        H, I = d.hidden_size, d.intermediate_size
        hq, hkv, D = d.num_attention_heads, d.num_key_value_heads, d.head_dim
        self.qkv_w = (hq + 2 * hkv) * D
        dev = self.device
        R = max_rows
        self.split_qkv = ops.choose_split_k(self.qkv_w, H, 32, self.sms)
        self.split_o = ops.choose_split_k(H, hq * D, 32, self.sms)
        self.split_down = ops.choose_split_k(H, I, 32, self.sms)
        # (tuning hooks: VB_SPLIT_QKV / VB_SPLIT_O / VB_SPLIT_DOWN override the split-K of the decode projections)
        self.split_qkv = int(os.environ.get("VB_SPLIT_QKV", self.split_qkv))
        self.split_o = int(os.environ.get("VB_SPLIT_O", self.split_o))
        self.split_down = int(os.environ.get("VB_SPLIT_DOWN", self.split_down))
        self._split_cache: Dict[Tuple[int, int], Tuple[int, int, int]] = {}
        # split-K partials: the largest (split x rows) any step size up to R needs (prefill-sized steps split as well, see
        # _splits: 24-40 weight tiles x one or two token tiles do not fill 148 SMs)
        need = max(max(self._splits(t)) * t for t in sorted({min(R, t) for t in (64, 256, 512, 768, 1024, R)}))
        self.hidden = torch.zeros(R, H, dtype=BF16, device=dev)
        self.normed = torch.zeros(R, H, dtype=BF16, device=dev)
        self.partials = torch.zeros(max(need, R) * max(self.qkv_w, H), dtype=torch.float32, device=dev)
        self.q = torch.zeros(R, hq, D, dtype=BF16, device=dev)
        self.attn = torch.zeros(R, hq, D, dtype=BF16, device=dev)
        self.act = torch.zeros(R, I, dtype=BF16, device=dev)
        # the projections read their activations in the tiled layout (one linear bulk copy per GEMM stage; a row-major
        # operand costs the TMA unit one request per token row -- 133 to 256 per stage in a prefill step, several times
        # the 16 KiB weight tile beside it): decode-sized AND prefill-sized steps (VB_PREFILL_TILED_ACTS=0: decode only)
        self.tiled_max_rows = R if os.environ.get("VB_PREFILL_TILED_ACTS", "1") != "0" else self.FUSED_MAX_ROWS
        self.normed_t = ops.TiledAct(self.FUSED_MAX_ROWS, H, dev, max_T=self.tiled_max_rows)
        self.attn_t = ops.TiledAct(self.FUSED_MAX_ROWS, hq * D, dev)
        self.act_t = ops.TiledAct(self.FUSED_MAX_ROWS, I, dev, max_T=self.tiled_max_rows)
        self.hidden_t = ops.TiledAct(self.FUSED_MAX_ROWS, H, dev)
        self.max_out_rows = min(R, 64)     # logits are only ever needed for one row per request
        self.last_normed = torch.zeros(self.max_out_rows, H, dtype=BF16, device=dev)
        self.logits = torch.zeros(self.max_out_rows, d.vocab_size, dtype=BF16, device=dev)
        self.attn_ws = ops.AttnWorkspace(R, hq, hkv, D, dev)
        self.freq = ops.rope_freq_table(d.rotary_dim or D, d.rope_factor, d.rope_theta, d.rope_interleave, d.low_freq_factor,
                                        d.high_freq_factor, d.old_context_len, device=dev)
        self.plan = ops.RowPlan(R, dev)
        self.attn_grid = self.attn_ws.grid
        # ---- fused decode path (<= FUSED_MAX_ROWS rows) ----
        self.gu_half = weights.gu_half
        # (the fused QKV tail has no q/k norm, no bias and rotates whole heads)
        self.fused_ok = (D in (64, 128) and H % 64 == 0 and not d.qk_norm and not d.qkv_bias and d.rotary_dim in (None, D)
                         and not d.rope_interleave)
        # split-K of the fused QKV projection: one head per tile
        self.fsplit_qkv = ops.proj_split_k(hq + 2 * hkv, H, self.sms)
        self.fsplit_o = ops.proj_split_k((H + 127) // 128, hq * D, self.sms)
        self.fsplit_down = ops.proj_split_k((H + 127) // 128, I, self.sms)
        tiles_h = (H + 127) // 128
        gu_tiles = (I + self.gu_half - 1) // self.gu_half
        self.ssq = torch.zeros(max(1, (H + 127) // 128) * self.FUSED_MAX_ROWS, dtype=torch.float32, device=dev)
        self.rope_cs = torch.zeros(self.FUSED_MAX_ROWS, 2, D, dtype=torch.float32, device=dev)

    FUSED_MAX_ROWS = 64

    def _partials(self, split: int, rows: int, width: int) -> torch.Tensor:
        return self.partials[: split * rows * width].view(split, rows, width)

    def _splits(self, rows: int) -> Tuple[int, int, int]:
        """split-K of the QKV / O / down projections of a step with ``rows`` token rows.  Decode-sized steps use the tuned
        values; larger steps re-apply the same rule to (weight tiles x 256-row token tiles): a 133-row prefill is ONE token
        tile, i.e. 24-40 CTAs per projection without split-K (measured: the 133-token Orpheus prefill took ~7 ms of a 63 ms
        TTFA that way)."""
        if rows <= self.FUSED_MAX_ROWS:
            return self.split_qkv, self.split_o, self.split_down
        if os.environ.get("VB_PREFILL_SPLIT", "1") == "0":           # (A/B switch: round-1 behaviour)
            return 1, 1, 1
        key = (rows + 255) // 256
        if key not in self._split_cache:
            d = self.dims
            H, I, hqD = d.hidden_size, d.intermediate_size, d.num_attention_heads * d.head_dim
            self._split_cache[key] = (ops.choose_split_k(self.qkv_w, H, rows, self.sms), ops.choose_split_k(H, hqD, rows, self.sms),
                                      ops.choose_split_k(H, I, rows, self.sms))
        return self._split_cache[key]

    def forward(self, input_ids: Optional[torch.Tensor], position_ids: torch.Tensor, n_rows: int,
                last_rows: Optional[torch.Tensor] = None, n_out: Optional[int] = None,
                plan: Optional[ops.RowPlan] = None, last_rows_offset: int = 0, head=None, want_hidden: bool = False):
        """input_ids / position_ids int32 [n_rows] on the device; self.plan must hold the step's row plan
        (ops.plan_rows).  Returns logits [n_out or n_rows, vocab] bf16 (a view of the static buffer).
        last_rows (int32 [n_out]) selects the rows whose logits are needed (prefill: qo_indptr[1:] - 1).
        input_ids None: the caller has already written the input embeddings into ``self.hidden[:n_rows]`` (models
        whose inputs are sums of several embeddings or projected features, csm.py:647-654).  ``head``: the packed
        output head to use instead of ``w.lm_head`` (per-position heads of a depth decoder).  ``want_hidden``: also
        return the final-normed rows [n_out or n_rows, hidden] bf16, row-major (the backbone state a depth decoder
        starts from, csm.py:290-300)."""
        d, w, R = self.dims, self.w, n_rows
        plan = self.plan if plan is None else plan
        if R > self.max_rows:
            raise VoxB200Error(f"{R} rows exceed the engine's max_rows {self.max_rows}")
        hidden, normed = self.hidden[:R], self.normed[:R]
        if input_ids is not None:
            ops.embedding(w.embed, input_ids, out=hidden)
        x_final = normed
        if R <= self.FUSED_MAX_ROWS and self.fused_ok and not self.force_unfused:
            self._layers_fused(position_ids, R, plan)
            if last_rows is None and self.tiled_acts:
                x_final = self.normed_t.view_rows(R)
            ops.rmsnorm(hidden, w.norm, d.rms_norm_eps, out=x_final)
        else:
            x_final = self._layers_unfused(position_ids, R, plan)
        hidden_rows = None
        if last_rows is not None or want_hidden:
            if isinstance(x_final, ops.TiledAct):      # a gather needs rows: redo the final norm row-major
                ops.rmsnorm(hidden, w.norm, d.rms_norm_eps, out=normed)
        if last_rows is not None:
            n_out = last_rows.numel() if n_out is None else n_out
            x = ops.gather_rows(normed, last_rows, out=self.last_normed[:n_out], idx_offset=last_rows_offset)
            hidden_rows = x
        else:
            n_out, x = R, x_final
            hidden_rows = normed
        if n_out > self.max_out_rows:
            raise VoxB200Error(f"logits requested for {n_out} rows; pass last_rows (max {self.max_out_rows})")
        logits = ops.gemm(x, w.lm_head if head is None else head, mode=0, out=self.logits[:n_out],
                          bias=w.head_bias if head is None else None)
        return (logits, hidden_rows) if want_hidden else logits

    # How decode-sized steps (<= FUSED_MAX_ROWS rows) run their layers.  Measured on B200 (Orpheus-3B, 32 rows): separate
    # kernels 2.61 ms per forward, the norm-/residual-fused projections 2.95 ms (VB_DECODE_MODE=fused; value 130 vs 142.7
    # audio-s/s, batch-1 TTFA 68 vs 62 ms): their cluster-reduced epilogues cost more than the launches they save
    # (profiles/r02_decode_modes.txt).  A third variant -- one persistent kernel per layer -- measured 3.2 ms and was
    # removed in round 2.  The fused projections stay as parity-tested C-ABI operators.
    force_unfused = os.environ.get("VB_DECODE_MODE", "unfused") == "unfused"

    def _layers_fused(self, position_ids: torch.Tensor, R: int, plan: ops.RowPlan) -> None:
        """hidden is updated in place, the RMSNorm statistics travel as per-tile sums of squares written by the
        residual projections (ssq[parts][R]).  Per layer: norm + QKV + RoPE + append, paged attention, O + residual,
        norm + gate/up + SiLU, down + residual: one launch per fused projection (5 per layer)."""
        d, w = self.dims, self.w
        hq, hkv, D, H, I = d.num_attention_heads, d.num_key_value_heads, d.head_dim, d.hidden_size, d.intermediate_size
        hidden, q, attn, act = self.hidden[:R], self.q[:R], self.attn[:R], self.act[:R]
        tiles_h = (H + 127) // 128
        ssq = self.ssq[: tiles_h * R].view(tiles_h, R)
        cs = ops.rope_table(position_ids[:R], self.freq, D, out=self.rope_cs[:R])
        ops.row_ssq(hidden, out=ssq[0])
        tiled = self.tiled_acts
        hidden_t = self.hidden_t.view_rows(R) if tiled else None
        attn_x_out = self.attn_t.view_rows(R) if tiled else attn          # what attention writes
        attn_x = attn_x_out if tiled else attn.view(R, hq * D)           # ... as the O projection reads it
        act_x = self.act_t.view_rows(R) if tiled else act
        eps = d.rms_norm_eps
        for i, L in enumerate(w.layers):
            # (layer 0 reads the embedding rows through the tensor map; later layers the tiled copy)
            ops.proj_norm_qkv_rope_append(hidden if (i == 0 or not tiled) else hidden_t, ssq, 1 if i == 0 else tiles_h,
                                          L["ln1"], eps, L["qkv"], self.kv_cache[i], cs, plan, hq, hkv, D,
                                          self.fsplit_qkv, q_out=q)
            ops.paged_attn(q, self.kv_map, i * self.pages_per_layer, plan, R, hkv, self.page_size, self.chunk,
                           self.attn_ws, out=attn_x_out, grid_ctas=self.attn_grid)
            ops.proj_residual(attn_x, L["o"], hidden, self.fsplit_o, hidden_out=hidden, ssq_out=ssq,
                              hidden_tiles_out=hidden_t)
            ops.proj_norm_gateup_silu(hidden_t if tiled else hidden, ssq, tiles_h, L["ln2"], eps, L["gu"], self.gu_half,
                                      I, out=act_x)
            ops.proj_residual(act_x, L["down"], hidden, self.fsplit_down, hidden_out=hidden, ssq_out=ssq,
                              hidden_tiles_out=hidden_t)

    def _layers_unfused(self, position_ids: torch.Tensor, R: int, plan: ops.RowPlan):
        """8 launches per layer.  Activations travel between kernels in the tiled layout (the attention output only in
        decode-sized steps and only with VB_ATTN_TILED=1); returns what holds the final normed rows (a TiledAct then,
        else self.normed[:R])."""
        d, w = self.dims, self.w
        hq, hkv, D, H, I = d.num_attention_heads, d.num_key_value_heads, d.head_dim, d.hidden_size, d.intermediate_size
        hidden = self.hidden[:R]
        tiled = R <= self.tiled_max_rows and self.tiled_acts
        normed = self.normed_t.view_rows(R) if tiled else self.normed[:R]
        attn_tiled = tiled and self.attn_tiled and R <= self.FUSED_MAX_ROWS
        attn_o = self.attn_t.view_rows(R) if attn_tiled else self.attn[:R]
        act = self.act_t.view_rows(R) if tiled else self.act[:R]
        ops.rmsnorm(hidden, w.layers[0]["ln1"], d.rms_norm_eps, out=normed)
        s_qkv, s_o, s_dn = self._splits(R)
        q = self.q[:R]
        n_layers = len(w.layers)
        for i, L in enumerate(w.layers):
            p = ops.gemm(normed, L["qkv"], mode=1, split_k=s_qkv, out=self._partials(s_qkv, R, self.qkv_w), tile_rows=D)
            ops.qkv_rope_append(p, self.kv_cache[i], position_ids, self.freq, plan, hq, hkv, D, q_out=q,
                                interleave=d.rope_interleave, q_norm=L.get("qn"), k_norm=L.get("kn"),
                                norm_eps=d.rms_norm_eps, qkv_bias=L.get("qkv_b"))
            ops.paged_attn(q, self.kv_map, i * self.pages_per_layer, plan, R, hkv, self.page_size, self.chunk,
                           self.attn_ws, out=attn_o, grid_ctas=self.attn_grid)
            p = ops.gemm(attn_o if attn_tiled else attn_o.view(R, hq * D), L["o"], mode=1, split_k=s_o,
                         out=self._partials(s_o, R, H))
            ops.reduce_residual_rmsnorm(p, hidden, L["ln2"], d.rms_norm_eps, hidden_out=hidden, normed_out=normed)
            ops.gemm(normed, L["gu"], mode=2, out=act, tile_rows=2 * self.gu_half, n_out=I)
            p = ops.gemm(act, L["down"], mode=1, split_k=s_dn, out=self._partials(s_dn, R, H))
            nxt = w.layers[i + 1]["ln1"] if i + 1 < n_layers else w.norm
            ops.reduce_residual_rmsnorm(p, hidden, nxt, d.rms_norm_eps, hidden_out=hidden, normed_out=normed)
        return normed

    tiled_acts = True         # tests: False = row-major activations + tensor-map loads in decode-sized steps too
    # attention output layout in the default mode: tiled (bulk-copy input of the O projection, but ~30 index
    # instructions per stored element in the attention kernel's result warps) or row-major (cheap stores, tensor-map
    # loads in the O projection); VB_ATTN_TILED=0/1
    attn_tiled = os.environ.get("VB_ATTN_TILED", "0") != "0"     # measured: row-major 2.645 ms vs tiled 2.678 ms per forward

    # ---- kernel-isolated passes for the roofline measurement (bench.py) --------------------------------
    def gemm_pass(self, n_rows: int) -> None:
        """Every projection launch of one decode step (QKV, O, gate/up, down per layer + lm_head) on the live
        buffers, nothing else: what bench.py replays to time the projection kernel alone.  Follows the engine's
        decode mode (separate kernels with tiled activations by default; fused projections otherwise)."""
        d, w, R = self.dims, self.w, n_rows
        hq, hkv, D, H, I = d.num_attention_heads, d.num_key_value_heads, d.head_dim, d.hidden_size, d.intermediate_size
        if self.force_unfused or not self.fused_ok:
            tiled = R <= self.FUSED_MAX_ROWS and self.tiled_acts
            normed = self.normed_t.view_rows(R) if tiled else self.normed[:R]
            attn_o = self.attn_t.view_rows(R) if tiled else self.attn[:R].view(R, hq * D)
            act = self.act_t.view_rows(R) if tiled else self.act[:R]
            s_qkv, s_o, s_dn = self._splits(R)
            for L in w.layers:
                ops.gemm(normed, L["qkv"], mode=1, split_k=s_qkv, out=self._partials(s_qkv, R, self.qkv_w), tile_rows=D)
                ops.gemm(attn_o, L["o"], mode=1, split_k=s_o, out=self._partials(s_o, R, H))
                ops.gemm(normed, L["gu"], mode=2, out=act, tile_rows=2 * self.gu_half, n_out=I)
                ops.gemm(act, L["down"], mode=1, split_k=s_dn, out=self._partials(s_dn, R, H))
            ops.gemm(normed, w.lm_head, mode=0, out=self.logits[:min(R, self.max_out_rows)])
            return
        hidden, q, attn, act = self.hidden[:R], self.q[:R], self.attn[:R], self.act[:R]
        tiles_h = (H + 127) // 128
        ssq = self.ssq[: tiles_h * R].view(tiles_h, R)
        cs = self.rope_cs[:R]
        for i, L in enumerate(w.layers):
            ops.proj_norm_qkv_rope_append(hidden, ssq, tiles_h, L["ln1"], d.rms_norm_eps, L["qkv"], self.kv_cache[i], cs,
                                          self.plan, hq, hkv, D, self.fsplit_qkv, q_out=q)
            ops.proj_residual(attn.view(R, hq * D), L["o"], hidden, self.fsplit_o, hidden_out=hidden,
                              ssq_out=ssq)
            ops.proj_norm_gateup_silu(hidden, ssq, tiles_h, L["ln2"], d.rms_norm_eps, L["gu"], self.gu_half, I, out=act)
            ops.proj_residual(act, L["down"], hidden, self.fsplit_down, hidden_out=hidden, ssq_out=ssq)
        ops.gemm(self.normed[:min(R, self.max_out_rows)], w.lm_head, mode=0, out=self.logits[:min(R, self.max_out_rows)])

    def gate_up_only(self, L: Dict[str, object], n_rows: int) -> None:
        """One gate/up projection launch (production decode mode) on the live buffers: bench.py's per-launch roofline."""
        d, R = self.dims, n_rows
        tiled = R <= self.FUSED_MAX_ROWS and self.tiled_acts
        normed = self.normed_t.view_rows(R) if tiled else self.normed[:R]
        act = self.act_t.view_rows(R) if tiled else self.act[:R]
        ops.gemm(normed, L["gu"], mode=2, out=act, tile_rows=2 * self.gu_half, n_out=d.intermediate_size)

    def attention_only(self, layer: int, n_rows: int, plan: ops.RowPlan) -> None:
        """The paged-attention launch of one layer on the live cache and the plan of the last step."""
        d = self.dims
        ops.paged_attn(self.q[:n_rows], self.kv_map, layer * self.pages_per_layer, plan, n_rows,
                       d.num_key_value_heads, self.page_size, self.chunk, self.attn_ws, out=self.attn[:n_rows],
                       grid_ctas=self.attn_grid)
