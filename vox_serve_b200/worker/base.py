"""B200 worker with the reference worker's public surface (``vox_serve/worker/base.py:14-775`` and
``worker/cuda_graph_worker.py:12-1280``): same constructor keywords, the five methods the schedulers call
(``prepare_lm_inputs``, ``run_lm_prefill``, ``run_lm_decode``, ``run_detokenize``, ``free_kv_cache``), the same
properties / attributes, and the same side effects on ``Request`` objects.

What is re-done underneath (SURVEY.md §3.2 "per step the host issues >= 4 device-wide synchronisations, one
FlashInfer plan, ~6 small H2D copies"):
  * every request owns a *batch slot*; its repetition cache row, last sampled id, token history ring and token
    counter live on the device for the request's lifetime (the reference re-stacks 5 MB of caches and
    concatenates 32 one-element tensors every step, worker/base.py:329-345);
  * one packed int32 staging buffer per step (page table, positions, slots) goes H2D in a single copy and the
    plan is derived on the device (vb_plan_rows) -- no host-side plan, no synchronize;
  * a decode step is ONE CUDA-graph replay: H2D staging copy -> plan -> input-id gather -> 28 layers -> lm_head
    -> fused sampler -> cache update -> feedback -> D2H of the sampled ids; graphs are captured per exact batch
    size on first use (no power-of-two padding with a temp page, cuda_graph_worker.py:966-977);
  * detokenize windows are gathered from the device-side history ring, SNAC + PCM16 run on the device and one
    D2H copy returns all chunks (the reference does one ``.cpu()`` per chunk, cuda_graph_worker.py:1252-1253).
There is no CPU / eager fallback: a missing libvoxb200.so raises at construction.
"""
from __future__ import annotations

import os
import queue
from typing import Coroutine, Dict, List, Optional

import numpy as np
import torch

from .. import _lib, ops
from .._lib import VoxB200Error
from ..flashinfer_utils import FlashInferDecodeWrapper, FlashInferPrefillWrapper
from ..model import load_model
from ..requests import LMInputs, Request

I32 = torch.int32


class _Staging:
    """Pinned host buffer + device twin holding one step's integer inputs; layout (int32):
    qo_indptr[B+1] | kv_indptr[B+1] | last_page_len[B] | slots[B] | position[T] | host_ids[T] | row_slot[T] |
    kv_indices[P].  Offsets are fixed by (max_req, max_rows) so CUDA graphs can bake the device pointers."""

    def __init__(self, max_req: int, max_rows: int, max_pages: int, device):
        self.max_req, self.max_rows, self.max_pages = max_req, max_rows, max_pages
        o = 0
        self.off = {}
        for name, n in (("qo", max_req + 1), ("indptr", max_req + 1), ("last", max_req), ("slots", max_req),
                        ("pos", max_rows), ("ids", max_rows), ("row_slot", max_rows), ("indices", max_pages)):
            self.off[name] = (o, n)
            o += (n + 3) // 4 * 4
        self.n = o
        self.host = torch.zeros(o, dtype=I32, pin_memory=True)
        self.np = self.host.numpy()
        self.dev = torch.zeros(o, dtype=I32, device=device)
        self.consumed = torch.cuda.Event()
        self.consumed.record()

    def h(self, name: str) -> np.ndarray:
        o, n = self.off[name]
        return self.np[o:o + n]

    def d(self, name: str, n: Optional[int] = None) -> torch.Tensor:
        o, cap = self.off[name]
        return self.dev[o:o + (cap if n is None else n)]

    def upload(self, n_ints: Optional[int] = None):
        n = self.n if n_ints is None else n_ints
        self.dev[:n].copy_(self.host[:n], non_blocking=True)
        self.consumed.record()

    def wait_consumed(self):
        """The host may refill the pinned buffer only after the previous step's H2D copy has executed."""
        self.consumed.synchronize()


class ModelWorker:
    _serves_depth = False

    def __init__(self, model_name: str, max_batch_size: int, max_num_pages: int, page_size: int, top_p: float = None,
                 top_k: int = None, min_p: float = None, temperature: float = None, max_tokens: int = None,
                 repetition_penalty: float = None, repetition_window: int = None, cfg_scale: float = None,
                 greedy: bool = False, enable_nvtx: bool = False, enable_torch_compile: bool = False,
                 detokenizer_device: Optional[str] = None, dp_rank: int = 0, dp_size: int = 1,
                 detokenize_interval: int = None, model=None, max_prefill_tokens: int = 1024, **model_kwargs):
        _lib.load()          # fail loudly here, not at the first step, if the CUDA library is absent
        if not torch.cuda.is_available():
            raise VoxB200Error("ModelWorker needs a CUDA device: there is no CPU path")
        if detokenizer_device not in (None, "cuda", "cuda:0", f"cuda:{torch.cuda.current_device()}"):
            raise VoxB200Error("LM/detokenizer disaggregation across GPUs is out of scope (SURVEY.md §8): replicas "
                               "are request-parallel, one full pipeline per GPU")
        self.device = f"cuda:{torch.cuda.current_device()}"
        self.detokenizer_device = self.device
        if model is None:
            model = load_model(model_name, device=self.device, top_p=top_p, top_k=top_k, min_p=min_p,
                               temperature=temperature, max_tokens=max_tokens, repetition_penalty=repetition_penalty,
                               repetition_window=repetition_window, cfg_scale=cfg_scale, greedy=greedy,
                               enable_torch_compile=enable_torch_compile, audio_decoder_device=self.device,
                               detokenize_interval=detokenize_interval, **model_kwargs)
        self.model = model
        mdev = torch.device(getattr(model, "device", self.device))
        if mdev.type == "cuda" and mdev.index is not None and mdev.index != torch.cuda.current_device():
            raise VoxB200Error(f"model lives on {mdev}, the worker on {self.device}: build the model on the replica's own "
                               "GPU (kernels are launched on the current device and read the weights directly)")
        if not 1 <= max_batch_size <= 64:
            raise VoxB200Error(f"max_batch_size {max_batch_size} outside [1, 64]: decode-sized steps (one row per request) are "
                               "what the projection kernels tile for; run more replicas for more streams")
        self.max_batch_size, self.dp_rank, self.dp_size = max_batch_size, dp_rank, dp_size
        self.max_num_pages, self.page_size = max_num_pages, page_size
        self.nvtx_enabled = enable_nvtx
        self.top_p, self.top_k, self.min_p, self.temperature = top_p, top_k, min_p, temperature
        self.repetition_penalty, self.repetition_window, self.cfg_scale = repetition_penalty, repetition_window, cfg_scale
        import logging

        self.logger = logging.getLogger(__name__)
        if self.model.has_depth_transformer and not self._serves_depth:
            # the reference builds ``CudaGraphWorker(model_name, ...)`` for every model (scheduler/base.py:57-77); the
            # multi-codebook step lives in a subclass here, selected once the model is known
            from .depth import DepthModelWorker

            self.__class__ = DepthModelWorker
        elif self._serves_depth and not self.model.has_depth_transformer:
            raise VoxB200Error("DepthModelWorker serves depth-transformer LMs only")
        # watermarking (silentcipher, worker/base.py:683-720) is a third-party post-filter outside the hot path: not applied
        self.needs_watermarking = False
        self.has_depth_transformer = self._serves_depth
        self.empty_pages: "queue.Queue[int]" = queue.Queue()
        for i in range(max_num_pages):
            self.empty_pages.put(i)
        # attributes the reference schedulers read (scheduler/base.py:243-245)
        self.prefill_graph_batch_size = max_batch_size      # no prefill-graph row limit here
        self.max_prefill_tokens = max_prefill_tokens
        self.cuda_graph_seq_len_buckets = [max_prefill_tokens]
        self._prepare_attention_wrappers()

    # ---- reference properties (worker/base.py:127-148) ------------------------------------------
    @property
    def detokenize_interval(self) -> int:
        return self.model.detokenize_interval

    @property
    def detokenize_overlap(self) -> int:
        return self.model.detokenize_overlap

    @property
    def supports_audio_input(self) -> bool:
        return self.model.supports_audio_input

    @property
    def available_batch_sizes(self) -> Optional[List[int]]:
        return None      # any batch size up to max_batch_size: graphs are captured per exact size

    # ---- device state ---------------------------------------------------------------------------
    def _prepare_attention_wrappers(self):
        m, dev, B = self.model, self.device, self.max_batch_size
        self.kv_cache = torch.zeros(m.num_hidden_layers, self.max_num_pages, 2, self.page_size, m.num_key_value_heads,
                                    m.head_dim, dtype=torch.bfloat16, device=dev)
        self.max_rows = self.max_prefill_tokens + B
        eng = m.engine_for(self.kv_cache, self.page_size)
        if eng.max_rows < self.max_rows:
            raise VoxB200Error(f"model engine holds {eng.max_rows} rows, worker needs {self.max_rows}")
        common = dict(attn_buffer=None, n_qo_head=m.num_attention_heads, n_kv_head=m.num_key_value_heads,
                      n_state=m.num_attention_heads * m.head_dim, page_size=self.page_size, device=dev,
                      max_pages=self.max_num_pages)
        self.prefill_wrapper = FlashInferPrefillWrapper(batch_size=B, max_seq_len=self.max_rows, **common)
        self.decode_wrapper = FlashInferDecodeWrapper(batch_size=B, use_cuda_graph=True, **common)
        self.staging = _Staging(B, self.max_rows, self.max_num_pages, dev)
        # slot-resident request state
        self.free_slots = list(range(B - 1, -1, -1))
        self.slot_of: Dict[str, int] = {}
        cfg = m.default_sampling_config
        self.rep_cache = None
        if m.use_repetition_penalty and cfg.repetition_window is not None:
            W = cfg.repetition_window if cfg.repetition_window > 0 else 1
            self.rep_cache = torch.zeros(B, W, m.n_codebooks, m.vocab_size, dtype=torch.bool, device=dev)
        self.history_cap = 1 << max(6, int(np.ceil(np.log2(m.max_tokens + 64))))
        self.next_input = torch.zeros(B, dtype=I32, device=dev)
        self.history = torch.zeros(B, self.history_cap, dtype=I32, device=dev)
        self.n_out = torch.zeros(B, dtype=I32, device=dev)
        self.input_ids = torch.zeros(self.max_rows, dtype=I32, device=dev)
        self.out_ids = torch.zeros(B, m.n_codebooks, dtype=torch.int64, device=dev)
        # sampled ids come back through a small ring of pinned buffers + events so that a scheduler running one
        # step ahead (scheduler/base.py:168-215) never reads a later step's ids
        self.ring = [(torch.zeros(B, m.n_codebooks, dtype=torch.int64, pin_memory=True), torch.cuda.Event())
                     for _ in range(4)]
        self.ring_pos = 0
        self.decode_graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.graph_pool = None
        self.use_cuda_graph = True
        # detokenize staging: [chunk] x (slot, first, n_valid)
        self.max_chunks = 4 * B
        self.win_host = torch.zeros(3, self.max_chunks, dtype=I32, pin_memory=True)
        self.win_dev = torch.zeros(3, self.max_chunks, dtype=I32, device=dev)
        self.pcm_host = torch.zeros(self.max_chunks, m.n_channels, m.output_audio_length, dtype=torch.int16,
                                    pin_memory=True)
        # the vocoder runs on its own stream: it only needs tokens the host has already seen, so it does not have to
        # queue behind the LM step that is still in flight when the scheduler runs one step ahead
        self.detok_stream = torch.cuda.Stream()
        # one CUDA graph per vocoder batch size: window gather -> frame de-interleave -> SNAC -> PCM16 as ONE replay
        # (the reference captures its detokenizer the same way, cuda_graph_worker.py:560-700); own memory pool because
        # these replays overlap LM steps on the main stream
        self.voc_graphs: Dict[int, tuple] = {}
        self.voc_pool = None
        self.voc_win = torch.zeros(self.max_chunks, m.detokenize_interval, dtype=torch.int64, device=dev)
        self.voc_pcm = torch.zeros(self.max_chunks, m.n_channels, m.output_audio_length, dtype=torch.int16, device=dev)
        self.gpu_launches = 0     # launches issued by this worker's own kernels (graph nodes counted at capture)
        self._graph_nodes: Dict[int, int] = {}

    def capture_decode_graphs(self, batch_sizes=None) -> int:
        """Capture the decode-step CUDA graph of every batch size up front, as the reference does at start-up
        (cuda_graph_worker.py:437-462 captures its batch-size buckets in __init__), so that no request pays a capture
        (~15-25 ms each) on its way to the first audio chunk.  Capture only records launches -- nothing executes, no
        request state is touched.  Returns the number of graphs captured."""
        n = 0
        for B in (batch_sizes or range(1, self.max_batch_size + 1)):
            if B not in self.decode_graphs:
                self._capture_decode(int(B))
                n += 1
            if self._vocoder_graphable() and B not in self.voc_graphs:
                self._capture_vocoder(int(B))
        return n

    def _vocoder_graphable(self) -> bool:
        dec = getattr(self.model, "audio_decoder", None)
        return self.use_cuda_graph and dec is not None and getattr(dec, "noise_source", None) is None

    def _vocoder_body(self, n: int):
        interval = self.detokenize_interval
        ops.gather_windows(self.history, self.win_dev[0], self.win_dev[1], self.win_dev[2], interval,
                           out=self.voc_win[:n], n=n)
        audio = self.model.postprocess(self.voc_win[:n].view(n, interval, 1))
        ops.pcm16(audio, out=self.voc_pcm[:n])

    def _capture_vocoder(self, n: int):
        dec = self.model.audio_decoder
        if hasattr(dec, "ensure_noise_state"):
            dec.ensure_noise_state()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        ds = self.detok_stream
        ds.wait_stream(torch.cuda.current_stream())
        before = ops.launch_count()
        with torch.cuda.stream(ds):
            with torch.cuda.graph(g, pool=self.voc_pool, stream=ds):
                self._vocoder_body(n)
        torch.cuda.current_stream().wait_stream(ds)
        if self.voc_pool is None:
            self.voc_pool = g.pool()
        self.voc_graphs[n] = (g, ops.launch_count() - before)
        return self.voc_graphs[n]

    # ---- step inputs (worker/base.py:210-360) ---------------------------------------------------
    def prepare_lm_inputs(self, lm_requests: List[Request], detokenize_requests: List[Request]) -> Optional[LMInputs]:
        for req in detokenize_requests:
            req.audio_decode_idx = req.next_audio_decode_idx.copy()
        if len(lm_requests) == 0:
            return None
        st, m = self.staging, self.model
        st.wait_consumed()
        qo, ip, last, slots = st.h("qo"), st.h("indptr"), st.h("last"), st.h("slots")
        pos, ids, row_slot, indices = st.h("pos"), st.h("ids"), st.h("row_slot"), st.h("indices")
        if len(lm_requests) > self.max_batch_size:
            raise VoxB200Error(f"{len(lm_requests)} LM requests exceed max_batch_size {self.max_batch_size}")
        is_prefill = any(not r.done_lm_prefill for r in lm_requests)
        failed: List[int] = []
        n_prefill_ok = 0
        t = 0         # rows so far
        npg = 0       # pages so far
        qo[0] = 0
        ip[0] = 0
        for i, req in enumerate(lm_requests):
            if not req.done_lm_prefill:
                if req.is_input_streaming and not m.supports_input_streaming:
                    raise ValueError(f"Input streaming is not supported by model {m.model_name}.")
                slot = self._acquire_slot(req)
                kw = dict(req.model_kwargs)
                if self.rep_cache is not None:
                    self.rep_cache[slot].zero_()
                    kw["repetition_cache"] = self.rep_cache[slot]
                out = m.preprocess(prompt=req.prompt, audio_path=req.audio_path, **kw)
                req.input_tokens = out.input_tokens
                n = int(req.input_tokens.shape[0])
                req.input_length = n
                if out.repetition_cache is not None:
                    req.repetition_cache = out.repetition_cache
                n_pages = (n + self.page_size - 1) // self.page_size
                # capacity: a request that cannot be served is finished with an error reason and everything it held
                # is released -- the scheduler loop (and the other streams) keep running.  (The reference raises out of
                # prepare_lm_inputs with half-updated state: queue.Empty / CUDA errors, worker/base.py:237-249.)
                why = None
                if n > self.max_prefill_tokens:
                    why = f"error: prompt of {n} tokens exceeds max_prefill_tokens {self.max_prefill_tokens}"
                elif t + n > self.max_rows - (len(lm_requests) - i - 1):
                    # fits a step of its own but not THIS one (several prompts selected at once: a scheduler that batches
                    # prefills does not know a prompt's length before preprocess has run): left un-prefilled for a later
                    # step, nothing held
                    s_ = self.slot_of.pop(req.request_id, None)
                    if s_ is not None:
                        self.free_slots.append(s_)
                    failed.append(i)
                    qo[i + 1], ip[i + 1], last[i], slots[i] = t, npg, 1, slot
                    continue
                elif self.empty_pages.qsize() < n_pages:
                    why = f"error: out of KV pages ({n_pages} needed, {self.empty_pages.qsize()} free)"
                if why is not None:
                    self._fail_request(req, why)
                    failed.append(i)
                    qo[i + 1], ip[i + 1], last[i], slots[i] = t, npg, 1, slot
                    continue
                self._stage_prompt_rows(req, out, t, n)
                pos[t:t + n] = np.arange(n, dtype=np.int32)
                row_slot[t:t + n] = -1
                req.kv_token_len = n
                req.kv_pages = [self.empty_pages.get_nowait() for _ in range(n_pages)]
                req.kv_last_page_len = n % self.page_size or self.page_size
                req.next_position_id = n + 1          # position n is skipped, as in worker/base.py:299
                req.done_lm_prefill = True
                n_prefill_ok += 1
                self.n_out[slot:slot + 1].zero_()
                t += n
            else:
                slot = self.slot_of[req.request_id]
                if req.kv_last_page_len + 1 > self.page_size and self.empty_pages.empty():
                    self._fail_request(req, "error: out of KV pages")
                    failed.append(i)
                    qo[i + 1], ip[i + 1], last[i], slots[i] = t, npg, 1, slot
                    continue
                req.kv_token_len += 1
                req.kv_last_page_len += 1
                if req.kv_last_page_len > self.page_size:
                    req.kv_pages.append(self.empty_pages.get_nowait())
                    req.kv_last_page_len = 1
                self._stage_decode_row(t)
                pos[t] = req.next_position_id
                row_slot[t] = slot
                req.next_position_id += 1
                t += 1
            k = len(req.kv_pages)
            indices[npg:npg + k] = req.kv_pages
            npg += k
            qo[i + 1] = t
            ip[i + 1] = npg
            last[i] = req.kv_last_page_len
            slots[i] = slot
        if failed:
            # drop the failed requests from the step (the caller's list is edited in place: it is what
            # run_lm_prefill / run_lm_decode receive) and rebuild the packed tables without their (empty) entries
            keep = [i for i in range(len(lm_requests)) if i not in failed]
            for dst, src in enumerate(keep):
                qo[dst + 1], ip[dst + 1], last[dst], slots[dst] = qo[src + 1], ip[src + 1], last[src], slots[src]
            lm_requests[:] = [lm_requests[i] for i in keep]
            if not lm_requests:
                return None
            is_prefill = n_prefill_ok > 0
        B = len(lm_requests)
        rep = None
        if self.rep_cache is not None:
            rep = self.rep_cache       # slot-resident; rows selected through `slots` (no per-step torch.stack)
        lm_inputs = {"qo_indptr": qo[:B + 1].tolist(), "paged_kv_indptr": ip[:B + 1].tolist(),
                     "paged_kv_indices": indices[:npg].tolist(),
                     "paged_kv_last_page_len": last[:B].tolist(), "input_ids": self._step_input_ids(t),
                     "position_ids": st.d("pos", t), "input_features": None, "input_masks": self._step_input_masks(t),
                     "repetition_cache": rep, "is_prefill": is_prefill, "n_rows": t, "n_pages": npg}
        if self.early_launch:
            lm_inputs["_launched"] = self._launch_step(B, t, is_prefill)
        return lm_inputs

    # ---- where a step's token rows come from (overridden by the multi-codebook worker) ----
    def _stage_prompt_rows(self, req: Request, out, t: int, n: int) -> None:
        tok = req.input_tokens[:, 0]
        self.staging.h("ids")[t:t + n] = (tok.cpu() if tok.is_cuda else tok).numpy()

    def _stage_decode_row(self, t: int) -> None:
        self.staging.h("ids")[t] = 0         # the row's id is the slot's last sampled token (build_input_ids)

    def _step_input_ids(self, t: int) -> torch.Tensor:
        return self.input_ids[:t].view(t, 1)

    def _step_input_masks(self, t: int) -> Optional[torch.Tensor]:
        return None

    def _fail_request(self, req: Request, reason: str) -> None:
        """Finish a request the worker cannot serve: no audio, an ``error: ...`` finish reason, slot and pages back in
        the pools.  The scheduler sends the completion message and drops it at its next step."""
        req.done_lm_prefill = req.done_lm_generation = req.done_all = True
        req.finish_reason = reason
        self.logger.error("request %s: %s", req.request_id, reason)
        self.free_kv_cache(req)

    def _acquire_slot(self, req: Request) -> int:
        if req.request_id in self.slot_of:
            return self.slot_of[req.request_id]
        if not self.free_slots:
            raise VoxB200Error("no free batch slot: more concurrent requests than max_batch_size")
        s = self.free_slots.pop()
        self.slot_of[req.request_id] = s
        return s

    # ---- LM steps -------------------------------------------------------------------------------
    def _device_step(self, B: int, T: int, is_prefill: bool):
        """Everything a step does on the device after the staging buffer is filled; capturable."""
        st, m = self.staging, self.model
        wrapper = self.prefill_wrapper if is_prefill else self.decode_wrapper
        wrapper.plan_device(st.d("qo", B + 1) if is_prefill else None, st.d("indptr", B + 1), st.d("indices"),
                            st.d("last", B), B, T)
        ops.build_input_ids(self.input_ids, st.d("ids"), self.next_input, st.d("row_slot"), T)
        last_rows = None
        if is_prefill:
            last_rows = st.d("qo", B + 1)[1:]         # qo_indptr[1:] (the kernel subtracts 1)
        logits = m.forward(self.input_ids[:T].view(T, 1), st.d("pos", T), wrapper, self.kv_cache,
                           logits_rows=last_rows, logits_rows_minus_one=is_prefill, n_rows=T)
        slots = st.d("slots", B)
        out = self.out_ids[:B]
        m.sampling_device(logits[:B], None, self.rep_cache, cache_rows=slots if self.rep_cache is not None else None,
                          out=out.view(-1))
        # the ring mirrors req.lm_output_audio_tokens index for index: the stop id is popped from that list on the
        # host (orpheus.py:461-463), so it is not recorded here either (a request that stopped may still run one more
        # LM step when the scheduler works one step ahead; its token then lands at the same index on both sides)
        ops.token_feedback(out.view(-1), slots, self.next_input, self.history, self.n_out,
                           skip_token=getattr(m, "stop_token_id", -1))

    # The schedulers call prepare_lm_inputs -> run_detokenize -> (send responses) -> run_lm_* (scheduler/base.py:149-162).
    # Everything the LM step needs is known when prepare_lm_inputs returns, so the step's device work is ENQUEUED there:
    # the vocoder pass of run_detokenize (side stream) and the host's wait for its PCM then overlap the LM step instead
    # of preceding it, also under the synchronous Scheduler._step.  run_lm_prefill / run_lm_decode hand back the
    # request-state coroutine of the step that is already in flight.  VB_EARLY_LAUNCH=0 restores launch-at-run_lm.
    early_launch = os.environ.get("VB_EARLY_LAUNCH", "1") != "0"

    def _launch_step(self, B: int, T: int, is_prefill: bool):
        self.nvtx_range_push(f"lm_{'prefill' if is_prefill else 'decode'}_bs{B}")
        self.staging.upload()
        if not is_prefill and self.use_cuda_graph:
            g = self.decode_graphs.get(B)
            if g is None:
                g = self._capture_decode(B)
            g.replay()
            self.gpu_launches += self._graph_nodes[B]
        else:
            self._device_step(B, T, is_prefill)
        ids_host, ready = self.ring[self.ring_pos]
        self.ring_pos = (self.ring_pos + 1) % len(self.ring)
        ids_host[:B].copy_(self.out_ids[:B], non_blocking=True)
        ready.record()
        self.nvtx_range_pop()
        return ids_host, ready, self.staging.h("slots")[:B].tolist()

    def _run_step(self, requests: List[Request], lm_inputs: LMInputs) -> Optional[Coroutine]:
        if len(requests) == 0:
            return None
        B, T, is_prefill = len(requests), lm_inputs["n_rows"], lm_inputs["is_prefill"]
        launched = lm_inputs.pop("_launched", None)
        ids_host, ready, slots_host = launched if launched is not None else self._launch_step(B, T, is_prefill)
        return self.model.sampling_host(self.out_ids[:B], requests, self.rep_cache, ids_host=ids_host, ready=ready,
                                        cache_rows_host=slots_host)

    def _capture_decode(self, B: int) -> torch.cuda.CUDAGraph:
        """Warm the kernels once eagerly on the live inputs is not possible (it would advance state twice), so the
        graph is captured directly: every launch inside is allocation-free and the library's one-time
        cudaFuncSetAttribute calls are legal during capture."""
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        before = ops.launch_count()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, pool=self.graph_pool, stream=s):
                self._device_step(B, B, False)
        torch.cuda.current_stream().wait_stream(s)
        if self.graph_pool is None:
            self.graph_pool = g.pool()
        self._graph_nodes[B] = ops.launch_count() - before
        self.decode_graphs[B] = g
        return g

    def run_lm_prefill(self, requests: List[Request], lm_inputs: LMInputs) -> Optional[Coroutine]:
        return self._run_step(requests, lm_inputs)

    def run_lm_decode(self, requests: List[Request], lm_inputs: LMInputs) -> Optional[Coroutine]:
        return self._run_step(requests, lm_inputs)

    # ---- detokenize (worker/base.py:616-681, cuda_graph_worker.py:1162-1280) ----------------------
    def run_detokenize(self, requests: List[Request]):
        if len(requests) == 0:
            return
        interval = self.detokenize_interval
        wh = self.win_host.numpy()
        mapping = []
        n = 0
        for ri, req in enumerate(requests):
            for ci, d in enumerate(req.audio_decode_idx):
                if n >= self.max_chunks:
                    raise VoxB200Error("more detokenize chunks in one call than the staging buffers hold")
                n_valid = min(interval, len(req.lm_output_audio_tokens) - d)
                if n_valid <= 0:
                    continue
                wh[0, n], wh[1, n], wh[2, n] = self.slot_of[req.request_id], d, n_valid
                mapping.append((ri, ci, n_valid))
                n += 1
        if n:
            self.nvtx_range_push(f"detokenize_bs{n}")
            ds = self.detok_stream
            # no cross-stream wait: every token a window refers to has already been read back by the host (its
            # step's ids-ready event was synchronised), so the history writes of those steps are complete
            graphable = self._vocoder_graphable()
            if graphable and n not in self.voc_graphs:
                self._capture_vocoder(n)             # (start-up normally captured 1..max_batch_size already)
            with torch.cuda.stream(ds):
                self.win_dev.copy_(self.win_host, non_blocking=True)
                if graphable:
                    g, nodes = self.voc_graphs[n]
                    g.replay()
                    self.gpu_launches += nodes
                else:                                # injected noise (parity tests): the eager launch sequence
                    self._vocoder_body(n)
                self.pcm_host[:n].copy_(self.voc_pcm[:n], non_blocking=True)
            ds.synchronize()
            self.nvtx_range_pop()
            pcm_np = self.pcm_host.numpy()
            for i, (ri, ci, n_valid) in enumerate(mapping):
                a16 = pcm_np[i]
                if n_valid < interval:      # drop the audio of the padded tokens (worker/base.py:661-668)
                    a16 = a16[:, :int(a16.shape[1] * (n_valid - 0.5) / interval)]
                requests[ri].output_audio.put(a16.tobytes())
        for req in requests:
            if req.done_lm_generation and req.audio_decode_idx and \
                    req.audio_decode_idx[-1] + interval >= len(req.lm_output_audio_tokens):
                req.done_all = True

    # ---- device-resident multi-step decode --------------------------------------------------------
    def run_lm_decode_resident(self, requests: List[Request], n_steps: int, detokenize: bool = True,
                               timing: Optional[dict] = None, tail_window: bool = False) -> int:
        """``n_steps`` decode steps for ``requests`` (all past prefill) with NO host work in between: pages for the
        whole span are allocated up front, kv lengths / positions / input ids / repetition caches / token history
        advance on the device, and each step is one CUDA-graph replay.  With ``detokenize`` the newest full
        window of every request is vocoded (SNAC + PCM16, result left in HBM) after every
        ``interval - overlap`` steps, as the scheduler would schedule it.  Afterwards the requests' host-side
        fields are brought up to date from the device history (stop ids are honoured only at the end of the span:
        tokens sampled after a stop id are dropped).  Returns the number of graph replays issued.

        This is the steady-state inner loop the per-step API converges to when nothing joins or leaves the batch;
        bench.py times it as the device-resident figure (``value``) next to the per-step API (``e2e``)."""
        B = len(requests)
        if B == 0 or n_steps <= 0:
            return 0
        if any(not r.done_lm_prefill or r.done_lm_generation for r in requests):
            raise VoxB200Error("resident decode needs requests that are past prefill and still generating")
        m, st = self.model, self.staging
        st.wait_consumed()
        ip, slots, indices, row_slot = st.h("indptr"), st.h("slots"), st.h("indices"), st.h("row_slot")
        kv0 = np.zeros(B, dtype=np.int32)
        pos0 = np.zeros(B, dtype=np.int32)
        npg = 0
        ip[0] = 0
        for i, req in enumerate(requests):
            need = (req.kv_token_len + n_steps + self.page_size - 1) // self.page_size
            while len(req.kv_pages) < need:
                req.kv_pages.append(self.empty_pages.get_nowait())
            k = len(req.kv_pages)
            indices[npg:npg + k] = req.kv_pages
            npg += k
            ip[i + 1] = npg
            slots[i] = row_slot[i] = self.slot_of[req.request_id]
            kv0[i], pos0[i] = req.kv_token_len, req.next_position_id - 1
        st.upload()
        if not hasattr(self, "res_kv_len"):
            self.res_kv_len = torch.zeros(self.max_batch_size, dtype=I32, device=self.device)
            self.res_pos = torch.zeros(self.max_batch_size, dtype=I32, device=self.device)
            self.res_first = torch.zeros(self.max_batch_size, dtype=I32, device=self.device)
            self.res_pcm = torch.zeros(self.max_batch_size, m.n_channels, m.output_audio_length, dtype=torch.int16,
                                       device=self.device)
            self.res_win = torch.zeros(self.max_batch_size, self.detokenize_interval, dtype=torch.int64,
                                       device=self.device)
            self.res_graphs: Dict[int, tuple] = {}
        self.res_kv_len[:B].copy_(torch.from_numpy(kv0), non_blocking=False)
        self.res_pos[:B].copy_(torch.from_numpy(pos0), non_blocking=False)

        def lm_step():
            ops.decode_advance(self.res_kv_len[:B], self.res_pos[:B])
            self.decode_wrapper.plan_device(None, st.d("indptr", B + 1), st.d("indices"), None, B, B,
                                            kv_len=self.res_kv_len)
            ops.build_input_ids(self.input_ids, st.d("ids"), self.next_input, st.d("row_slot"), B)
            logits = m.forward(self.input_ids[:B].view(B, 1), self.res_pos[:B], self.decode_wrapper, self.kv_cache,
                               n_rows=B)
            sl = st.d("slots", B)
            out = self.out_ids[:B]
            m.sampling_device(logits[:B], None, self.rep_cache, cache_rows=sl if self.rep_cache is not None else None,
                              out=out.view(-1))
            ops.token_feedback(out.view(-1), sl, self.next_input, self.history, self.n_out)

        # The vocoder runs on the side stream (as in run_detokenize), overlapped with the following LM steps.  What
        # must stay ordered with the LM steps is only the selection of the window: token_feedback of the next step
        # advances n_out / history, so latest_window + gather_windows are replayed in line on the main stream and the
        # SNAC graph reads the gathered copy.
        def window_step():
            sl = st.d("slots", B)
            ops.latest_window(self.res_first, self.n_out, sl, B, self.detokenize_interval)
            ops.gather_windows(self.history, sl, self.res_first, None, self.detokenize_interval, out=self.res_win[:B],
                               n=B)

        def vocoder_step():
            audio = m.postprocess(self.res_win[:B].view(B, self.detokenize_interval, 1))
            ops.pcm16(audio, out=self.res_pcm[:B])

        if B not in self.res_graphs:
            graphs = []
            for fn in (lm_step, window_step, vocoder_step):
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                before = ops.launch_count()
                with torch.cuda.stream(s):
                    with torch.cuda.graph(g, pool=self.graph_pool, stream=s):
                        fn()
                torch.cuda.current_stream().wait_stream(s)
                if self.graph_pool is None:
                    self.graph_pool = g.pool()
                graphs.append((g, ops.launch_count() - before))
            self.res_graphs[B] = tuple(graphs)
        (g_lm, n_lm), (g_win, n_win), (g_voc, n_voc) = self.res_graphs[B]
        hop = self.detokenize_interval - self.detokenize_overlap
        replays = 0
        main, side = torch.cuda.current_stream(), self.detok_stream
        overlap = os.environ.get("VB_RESIDENT_VOCODER_OVERLAP", "1") != "0"
        side.wait_stream(main)
        if timing is not None:
            timing["start"].record()
        for k in range(n_steps):
            g_lm.replay()
            self.gpu_launches += n_lm
            replays += 1
            if detokenize and ((k + 1) % hop == 0 or (tail_window and k + 1 == n_steps)):
                # (tail_window: a span that is not a whole number of hops still ends with a vocoder pass, so that
                # a timed span accounts for at least the vocoder work its tokens need)
                main.wait_stream(side)           # the previous vocoder pass has consumed res_win
                g_win.replay()
                if overlap:
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        g_voc.replay()
                else:
                    g_voc.replay()
                self.gpu_launches += n_win + n_voc
                replays += 2
        main.wait_stream(side)
        if timing is not None:
            timing["end"].record()
            timing["lm_nodes"], timing["detok_nodes"] = n_lm, n_win + n_voc
        # ---- bring the host-side request state up to date ----
        torch.cuda.synchronize()
        hist = self.history.cpu()
        n_now = self.n_out.cpu()
        for i, req in enumerate(requests):
            s_ = int(slots[i])
            n_new = n_steps
            base = int(n_now[s_]) - n_new
            toks = [int(hist[s_, (base + j) % self.history_cap]) for j in range(n_new)]
            for j, tok in enumerate(toks):
                t = torch.tensor([[tok]], dtype=torch.int64)
                req.lm_output_tokens.append(t)
                req.lm_output_audio_tokens.append(t)
                req.kv_token_len += 1
                req.next_position_id += 1
                if tok == m.stop_token_id:
                    req.lm_output_audio_tokens.pop()
                    req.done_lm_generation, req.finish_reason = True, "stop_id_encountered"
                    break
                if req.next_position_id > m.max_tokens:
                    req.done_lm_generation, req.finish_reason = True, "max_tokens_reached"
                    break
            req.kv_last_page_len = req.kv_token_len % self.page_size or self.page_size
            keep = (req.kv_token_len + self.page_size - 1) // self.page_size
            while len(req.kv_pages) > keep:          # pages reserved for the span but not reached
                self.empty_pages.put(req.kv_pages.pop())
            req.input_tokens = self.out_ids[i:i + 1]
        return replays

    # ---- misc ---------------------------------------------------------------------------------------
    def free_kv_cache(self, request: Request):
        if getattr(request, "kv_pages", None):
            for p in request.kv_pages:
                self.empty_pages.put(p)
            request.kv_pages = []
            request.kv_token_len = 0
            request.kv_last_page_len = 0
        s = self.slot_of.pop(request.request_id, None)
        if s is not None:
            self.free_slots.append(s)

    def nvtx_range_push(self, name: str):
        if self.nvtx_enabled:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_push(name)

    def nvtx_range_pop(self):
        if self.nvtx_enabled:
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()


class CudaGraphWorker(ModelWorker):
    """Name kept for ``isinstance(worker, CudaGraphWorker)`` checks (scheduler/base.py:242): the B200 worker
    always replays decode steps from CUDA graphs."""
