"""Oracle restatement of the worker / scheduler bookkeeping around the hot path.

* page allocation, positions, LMInputs lists: ``vox_serve/worker/base.py:210-360``
  (first decode position is ``prompt_len + 1`` -- ``base.py:299`` -- reproduced as is)
* last-token logits gather for prefill: ``vox_serve/worker/cuda_graph_worker.py:900-902``
  (``qo_indptr[1:] - 1``; the non-graph worker's ``qo_indptr[:-1] - 1`` is a reference bug)
* request state update after sampling: ``vox_serve/model/orpheus.py:449-473``
* detokenize window selection: ``vox_serve/scheduler/base.py:302-333``
* window padding / trim / PCM16 / done_all: ``vox_serve/worker/cuda_graph_worker.py:1176-1277``
  (``(audio * 32767).astype(int16)``: truncation toward zero, no clipping)
* LM request selection (one prefill per step + up to 7 riding decodes): ``scheduler/base.py:234-300``

Drives ``oracle.orpheus`` + ``oracle.snac`` on CPU.  Test infrastructure only.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import lm_ops, orpheus as oorph, sampler as osampler, snac as osnac


@dataclass
class Req:
    request_id: str
    prompt_ids: torch.Tensor                      # [T] int64 (already formatted, orpheus.py:356-358)
    next_position_id: Optional[int] = None
    kv_pages: List[int] = field(default_factory=list)
    kv_token_len: int = 0
    kv_last_page_len: int = 0
    input_tokens: Optional[torch.Tensor] = None   # [T, 1]
    lm_output_tokens: List[int] = field(default_factory=list)
    lm_output_audio_tokens: List[int] = field(default_factory=list)
    repetition_cache: Optional[torch.Tensor] = None
    done_lm_prefill: bool = False
    done_lm_generation: bool = False
    done_all: bool = False
    finish_reason: Optional[str] = None
    audio_decode_idx: List[int] = field(default_factory=list)
    next_audio_decode_idx: List[int] = field(default_factory=list)
    output_audio: List[bytes] = field(default_factory=list)


def format_prompt(text_ids: Sequence[int]) -> torch.Tensor:
    """[128259] + ids + [128009, 128260, 128261, 128257]  (orpheus.py:356-358)."""
    return torch.tensor([128259, *text_ids, 128009, 128260, 128261, 128257], dtype=torch.int64)


def pcm16(audio: np.ndarray) -> np.ndarray:
    return (audio * 32767).astype(np.int16)


def trim_len(n_samples: int, last_chunk_len: int, interval: int) -> int:
    return int(n_samples * (last_chunk_len - 0.5) / interval)


class OracleWorker:
    def __init__(self, weights, dims: oorph.OrpheusDims, cfg, page_size: int = 128, max_num_pages: int = 2048,
                 snac_sd=None, snac_cfg: Optional[osnac.SnacConfig] = None, max_batch_size: int = 32,
                 detokenize_interval: int = 28, detokenize_overlap: int = 21, noise_seed: int = 1234,
                 ignore_stop: bool = False):
        self.w, self.dims, self.cfg = weights, dims, cfg
        self.page_size, self.max_num_pages = page_size, max_num_pages
        self.empty_pages = deque(range(max_num_pages))
        dt = weights["lm_head.weight"].dtype
        self.kv_cache = torch.zeros(dims.num_hidden_layers, max_num_pages, 2, page_size,
                                    dims.num_key_value_heads, dims.head_dim, dtype=dt)
        self.snac_sd, self.snac_cfg = snac_sd, snac_cfg
        self.max_batch_size = max_batch_size
        self.interval, self.overlap = detokenize_interval, detokenize_overlap
        self.noise_gen = torch.Generator().manual_seed(noise_seed)
        self.ignore_stop = ignore_stop
        self.use_rep = (cfg.repetition_penalty is not None and cfg.repetition_window is not None
                        and cfg.repetition_penalty != 1.0)
        self.last_logits = None
        self.last_penalised = None
        self.last_own_ids = None

    # ---- worker/base.py:210-360 ----
    def prepare_lm_inputs(self, reqs: List[Req]) -> Optional[Dict]:
        if not reqs:
            return None
        qo, ip, idx, last, ids, pos = [0], [0], [], [], [], []
        is_prefill = any(not r.done_lm_prefill for r in reqs)
        for r in reqs:
            if not r.done_lm_prefill:
                r.input_tokens = r.prompt_ids.view(-1, 1)
                n = r.input_tokens.shape[0]
                if self.use_rep:
                    w = self.cfg.repetition_window if self.cfg.repetition_window > 0 else 1
                    r.repetition_cache = torch.zeros(w, 1, self.dims.vocab_size, dtype=torch.bool)
                ids.append(r.input_tokens)
                pos.extend(range(n))
                r.kv_token_len = n
                r.kv_pages = [self.empty_pages.popleft() for _ in range((n + self.page_size - 1) // self.page_size)]
                r.kv_last_page_len = n % self.page_size or self.page_size
                qo.append(qo[-1] + n)
                r.next_position_id = n + 1
                r.done_lm_prefill = True
            else:
                ids.append(r.input_tokens)
                r.kv_token_len += 1
                r.kv_last_page_len += 1
                if r.kv_last_page_len > self.page_size:
                    r.kv_pages.append(self.empty_pages.popleft())
                    r.kv_last_page_len = 1
                qo.append(qo[-1] + 1)
                pos.append(r.next_position_id)
                r.next_position_id += 1
            ip.append(ip[-1] + len(r.kv_pages))
            idx.extend(r.kv_pages)
            last.append(r.kv_last_page_len)
        rep = torch.stack([r.repetition_cache for r in reqs], 0) if self.use_rep else None
        return dict(qo_indptr=qo, paged_kv_indptr=ip, paged_kv_indices=idx, paged_kv_last_page_len=last,
                    input_ids=torch.cat(ids, 0), position_ids=torch.tensor(pos, dtype=torch.int32),
                    repetition_cache=rep, is_prefill=is_prefill)

    # ---- run_lm_prefill / run_lm_decode + orpheus.py:419-477 ----
    def run_lm(self, reqs: List[Req], inp: Optional[Dict], forced_ids: Optional[torch.Tensor] = None
               ) -> Optional[torch.Tensor]:
        """``forced_ids`` ([B, 1] int64): teacher forcing for parity tests -- the oracle's own choice is computed
        (returned, with its penalised logits kept in ``last_penalised``) but the request state and the repetition
        cache advance with the forced ids, so one bf16 near-tie does not fork the rest of the comparison."""
        if not reqs:
            return None
        if inp["is_prefill"]:
            wr = lm_ops.PagedWrapperCPU("prefill", self.page_size)
            wr.plan(inp["qo_indptr"], inp["paged_kv_indptr"], inp["paged_kv_indices"], inp["paged_kv_last_page_len"])
        else:
            wr = lm_ops.PagedWrapperCPU("decode", self.page_size)
            wr.plan(inp["paged_kv_indptr"], inp["paged_kv_indices"], inp["paged_kv_last_page_len"])
        logits = oorph.lm_forward(self.w, self.dims, inp["input_ids"][:, 0], inp["position_ids"], wr, self.kv_cache)
        if inp["is_prefill"]:
            logits = logits[torch.tensor(inp["qo_indptr"][1:]) - 1]
        logits = logits[:, None, :]
        self.last_logits = logits
        if self.ignore_stop:
            logits = logits.clone()
            logits[..., self.dims.stop_token_id] = float("-inf")
        rep = inp["repetition_cache"]
        if forced_ids is not None:
            self.last_penalised = (osampler.apply_repetition_penalty(logits, rep, self.cfg.repetition_penalty)
                                   if rep is not None else logits)
            own = oorph.sampling_step(logits, self.cfg, rep.clone() if rep is not None else None)
            if rep is not None:
                osampler.update_repetition_cache(rep, forced_ids, self.cfg.repetition_window)
            ids, self.last_own_ids = forced_ids, own
        else:
            ids = oorph.sampling_step(logits, self.cfg, rep)
        for i, r in enumerate(reqs):
            tok = int(ids[i, 0])
            r.input_tokens = ids[i : i + 1]
            r.lm_output_tokens.append(tok)
            r.lm_output_audio_tokens.append(tok)
            if tok == self.dims.stop_token_id:
                r.lm_output_audio_tokens.pop()
                r.done_lm_generation, r.finish_reason = True, "stop_id_encountered"
        for i, r in enumerate(reqs):
            if r.next_position_id > self.dims.max_tokens:
                r.done_lm_generation, r.finish_reason = True, "max_tokens_reached"
            if self.use_rep:
                r.repetition_cache = inp["repetition_cache"][i]
        return ids

    # ---- scheduler/base.py:302-333 ----
    def select_detokenize(self, active: List[Req]) -> List[Req]:
        out, step = [], self.interval - self.overlap
        for r in active:
            if len(out) >= self.max_batch_size:
                break
            nxt = r.next_audio_decode_idx[-1] + step if r.next_audio_decode_idx else 0
            if r.done_lm_generation:
                if nxt < len(r.lm_output_audio_tokens):
                    r.next_audio_decode_idx = [nxt]
                else:
                    r.done_all = True
                out.append(r)
            elif nxt + self.interval <= len(r.lm_output_audio_tokens):
                r.next_audio_decode_idx = [nxt]
                out.append(r)
        return out

    # ---- scheduler/base.py:234-300 ----
    def select_lm(self, active: List[Req], prefill_graph_batch_size: int = 8) -> List[Req]:
        pre = [r for r in active if not r.done_lm_generation and not r.done_lm_prefill]
        dec = [r for r in active if not r.done_lm_generation and r.done_lm_prefill]
        out: List[Req] = []
        if pre:
            out.append(pre[0])
            slots = prefill_graph_batch_size - 1
        else:
            slots = self.max_batch_size
        for r in dec[:slots]:
            if len(out) >= self.max_batch_size:
                break
            out.append(r)
        return out

    # ---- cuda_graph_worker.py:1162-1280 + orpheus.py:483-507 ----
    def windows(self, reqs: List[Req]):
        toks, mapping = [], []
        for ri, r in enumerate(reqs):
            for ci, d in enumerate(r.audio_decode_idx):
                t = list(r.lm_output_audio_tokens[d : d + self.interval])
                if len(t) < self.interval:
                    t.extend([t[-1]] * (self.interval - len(t)))
                toks.append(t)
                mapping.append((ri, ci))
        return toks, mapping

    def run_detokenize(self, reqs: List[Req], noises=None):
        for r in reqs:  # worker/base.py:216-217
            r.audio_decode_idx = list(r.next_audio_decode_idx)
        toks, mapping = self.windows(reqs)
        if not toks:
            return None
        ids = torch.tensor(toks, dtype=torch.int64).view(len(toks), self.interval, 1)
        codes = oorph.audio_codes_from_window(ids, self.dims)
        if noises is None:
            noises = [torch.randn(s, generator=self.noise_gen) for s in
                      osnac.noise_shapes(self.snac_cfg, len(toks), codes[-1].shape[1])]
        wav = osnac.decode(self.snac_sd, self.snac_cfg, codes, noises)
        n_total = wav.shape[-1]
        audio = wav[:, :, n_total // 4 : n_total // 2]      # [2048:4096] of 8192 (orpheus.py:506)
        for i, (ri, ci) in enumerate(mapping):
            r = reqs[ri]
            d = r.audio_decode_idx[ci]
            a16 = pcm16(audio[i].numpy())
            n_last = len(r.lm_output_audio_tokens[d : d + self.interval])
            if n_last < self.interval:
                a16 = a16[:, : trim_len(a16.shape[1], n_last, self.interval)]
            r.output_audio.append(a16.tobytes())
        for r in reqs:
            if r.done_lm_generation and r.audio_decode_idx and \
                    r.audio_decode_idx[-1] + self.interval >= len(r.lm_output_audio_tokens):
                r.done_all = True
        return audio

    def free_kv_cache(self, r: Req):
        self.empty_pages.extend(r.kv_pages)
        r.kv_pages, r.kv_token_len, r.kv_last_page_len = [], 0, 0

    # ---- Scheduler._step, scheduler/base.py:135-166 (transport removed) ----
    def step(self, active: List[Req]):
        det = self.select_detokenize(active)
        lm = self.select_lm(active)
        inp = self.prepare_lm_inputs(lm)
        if self.snac_sd is not None:
            self.run_detokenize(det)
        else:
            for r in det:
                r.audio_decode_idx = list(r.next_audio_decode_idx)
        for r in det:
            if r.done_all:
                self.free_kv_cache(r)
        self.run_lm(lm, inp)
        return lm, det
