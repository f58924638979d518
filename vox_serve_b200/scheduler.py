"""In-process scheduler loop: ``Scheduler._step`` of the reference (``vox_serve/scheduler/base.py:135-166``) with
its ZMQ transport replaced by Python queues -- request intake, detokenize / LM selection policies
(``:234-333``), the fixed call order into the five worker methods, response collection and page release
(``:335-363``).  It exists so the hot path can be driven, timed (bench.py) and parity-tested end to end without
the HTTP / ZMQ control plane, which is out of scope; the reference's own schedulers
drive the same worker methods unchanged.
"""
from __future__ import annotations

import time
from collections import deque
from typing import Callable, Deque, Dict, List, Optional

from .requests import Request


class Scheduler:
    def __init__(self, worker, max_batch_size: Optional[int] = None, one_prefill_per_step: bool = True,
                 on_audio: Optional[Callable[[Request, bytes, float], None]] = None, vocoder_batch_steps: int = 1):
        self.model_worker = worker
        # 1 = the reference's schedule (every request is vocoded in the step its window completes).  N > 1 holds
        # completed windows back until every N-th step so that the vocoder runs on a full batch: with staggered
        # requests the per-step batches are ~batch/7 windows and the early decoder stages cost as much for 5 windows
        # as for 32.  A request's FIRST chunk and the chunks of finished requests are never held (TTFA unchanged);
        # later chunks arrive at most N - 1 steps late (a chunk is 85 ms of audio, a step ~3 ms).
        self.vocoder_batch_steps = max(1, int(vocoder_batch_steps))
        self._select_no = 0
        self.max_batch_size = max_batch_size or worker.max_batch_size
        self.one_prefill_per_step = one_prefill_per_step
        self.pending: Deque[Request] = deque()
        self.active_requests: List[Request] = []
        self.finished: List[Request] = []
        self.audio: Dict[str, List[bytes]] = {}
        self.first_audio_time: Dict[str, float] = {}
        self.submit_time: Dict[str, float] = {}
        self.on_audio = on_audio
        self.steps = 0
        self.trace = None      # set to [] to record [(request_id, sampled id)] per step (parity tests)

    # ---- intake (scheduler/base.py:431-478 without the socket) ------------------------------------
    def submit(self, req: Request):
        self.submit_time[req.request_id] = time.perf_counter()
        self.audio[req.request_id] = []
        self.pending.append(req)

    def detach(self, request_id: str) -> Request:
        """Take a request out of this loop WITHOUT finishing it (it is being handed to another replica,
        vox_serve_b200/kv_handoff.py); the caller releases or transfers what the worker holds for it."""
        for pool in (self.active_requests, self.pending):
            for r in list(pool):
                if r.request_id == request_id:
                    pool.remove(r)
                    return r
        raise KeyError(request_id)

    def adopt(self, req: Request) -> None:
        """Continue a request that arrived through a KV hand-off: its prefill is done and the worker already holds its
        pages, batch slot and decode state, so it joins the active set directly."""
        self.submit_time.setdefault(req.request_id, time.perf_counter())
        self.audio.setdefault(req.request_id, [])
        self.active_requests.append(req)

    def _prepare_requests(self):
        for r in self.active_requests:
            # a request the worker FAILED (finish_reason "error: ...": prompt too long, out of KV pages) never reaches
            # _send_responses: its completion is recorded here (the worker has already released what it held)
            if r.done_all and (r.finish_reason or "").startswith("error") and r not in self.finished:
                self.finished.append(r)
        self.active_requests = [r for r in self.active_requests if not r.done_all]
        while self.pending and len(self.active_requests) < self.max_batch_size:
            self.active_requests.append(self.pending.popleft())

    # ---- selection policies -----------------------------------------------------------------------
    def _select_detokenize_requests(self) -> List[Request]:
        """scheduler/base.py:302-333."""
        out: List[Request] = []
        interval, overlap = self.model_worker.detokenize_interval, self.model_worker.detokenize_overlap
        step = interval - overlap
        self._select_no += 1
        hold = self.vocoder_batch_steps > 1 and self._select_no % self.vocoder_batch_steps != 0
        for req in self.active_requests:
            if len(out) >= self.max_batch_size:
                break
            if req.done_all:
                # only reachable when running one step ahead (_step_async): the request finished in the step that
                # was just issued and is dropped at the next _prepare_requests; selecting it again would vocode
                # its last window twice
                continue
            nxt = req.next_audio_decode_idx[-1] + step if req.next_audio_decode_idx else 0
            if req.done_lm_generation:
                if nxt < len(req.lm_output_audio_tokens):
                    req.next_audio_decode_idx = [nxt]
                else:
                    req.done_all = True
                out.append(req)
            elif nxt + interval <= len(req.lm_output_audio_tokens):
                if hold and req.next_audio_decode_idx:
                    continue
                req.next_audio_decode_idx = [nxt]
                out.append(req)
        return out

    def _select_lm_requests(self) -> List[Request]:
        """scheduler/base.py:234-300: at most one prefill per step, decode requests fill the remaining rows."""
        w = self.model_worker
        max_prefill_bs = getattr(w, "prefill_graph_batch_size", self.max_batch_size)
        max_seq_len = max(getattr(w, "cuda_graph_seq_len_buckets", [1024]))
        prefill = [r for r in self.active_requests if not r.done_lm_generation and not r.done_lm_prefill]
        decode = [r for r in self.active_requests if not r.done_lm_generation and r.done_lm_prefill]
        out: List[Request] = []
        if prefill:
            seq = 0
            for r in prefill:
                n = r.input_length or 0
                if len(out) + 1 <= max_prefill_bs and seq + n <= max_seq_len:
                    out.append(r)
                    seq += n
                if self.one_prefill_per_step or len(out) >= max_prefill_bs:
                    break
            remaining = max_prefill_bs - len(out)
        else:
            remaining = self.max_batch_size
        for r in decode[:remaining]:
            if len(out) >= self.max_batch_size:
                break
            out.append(r)
        return out

    # ---- responses (scheduler/base.py:335-363) ------------------------------------------------------
    def _send_responses(self, detokenize_requests: List[Request]):
        now = time.perf_counter()
        for req in detokenize_requests:
            while not req.output_audio.empty():
                chunk = req.output_audio.get()
                self.audio[req.request_id].append(chunk)
                self.first_audio_time.setdefault(req.request_id, now)
                if self.on_audio is not None:
                    self.on_audio(req, chunk, now)
            if req.done_all:
                self.model_worker.free_kv_cache(req)
                self.finished.append(req)

    # ---- one iteration: the hot loop ----------------------------------------------------------------
    def _step(self):
        self._prepare_requests()
        detokenize_requests = self._select_detokenize_requests()
        lm_requests = self._select_lm_requests()
        w = self.model_worker
        lm_inputs = w.prepare_lm_inputs(lm_requests, detokenize_requests)
        w.run_detokenize(detokenize_requests)
        self._send_responses(detokenize_requests)
        if lm_inputs is not None and lm_inputs["is_prefill"]:
            task = w.run_lm_prefill(lm_requests, lm_inputs)
        else:
            task = w.run_lm_decode(lm_requests, lm_inputs)
        if task is not None:
            # scheduler/base.py:164-165 runs the state-update coroutine with asyncio.run(); it never suspends on
            # anything but the device event, so it is driven directly (no event-loop construction per step)
            while True:
                try:
                    task.send(None)
                except StopIteration:
                    break
        if self.trace is not None:
            self.trace.append([(r.request_id, int(r.lm_output_tokens[-1][0, 0])) for r in lm_requests])
        self.steps += 1
        return len(lm_requests), len(detokenize_requests)

    def _step_async(self, task, lm_requests: List[Request], detokenize_requests: List[Request]):
        """``Scheduler._step_async`` (scheduler/base.py:168-215): launch this step's model work first, THEN finish
        the previous step's request-state update and select the next step's requests, so the host's bookkeeping
        overlaps the device.  Request state therefore lags one step behind the device, exactly as in the
        reference's ``--async-scheduling`` mode."""
        self._prepare_requests()
        w = self.model_worker
        lm_inputs = w.prepare_lm_inputs(lm_requests, detokenize_requests)
        # ---- run_model ----
        w.run_detokenize(detokenize_requests)
        self._send_responses(detokenize_requests)
        if lm_inputs is not None and lm_inputs["is_prefill"]:
            next_task = w.run_lm_prefill(lm_requests, lm_inputs)
        else:
            next_task = w.run_lm_decode(lm_requests, lm_inputs)
        # ---- run_scheduling ----
        if task is not None:
            while True:
                try:
                    task.send(None)
                except StopIteration:
                    break
        if self.trace is not None:
            self.trace.append([(r.request_id, None) for r in lm_requests])
        next_detokenize_requests = self._select_detokenize_requests()
        next_lm_requests = self._select_lm_requests()
        self.steps += 1
        return next_task, next_lm_requests, next_detokenize_requests

    def run_async(self, n_steps: Optional[int] = None, state=None):
        """Drive ``_step_async`` (scheduler/base.py:217-221); returns the carried (task, lm, detokenize) state so
        the loop can be continued."""
        task, lm, det = state if state is not None else (None, [], [])
        n = 0
        while (n_steps is None and (self.has_work() or task is not None)) or (n_steps is not None and n < n_steps):
            task, lm, det = self._step_async(task, lm, det)
            n += 1
        return task, lm, det

    def has_work(self) -> bool:
        return bool(self.pending) or any(not r.done_all for r in self.active_requests)

    def run_until_done(self, max_steps: int = 1 << 30):
        n = 0
        while self.has_work() and n < max_steps:
            self._step()
            n += 1
        self._prepare_requests()
        return n

    def audio_seconds(self, sample_rate: int = 24000, bytes_per_sample: int = 2) -> float:
        return sum(len(c) for chunks in self.audio.values() for c in chunks) / (sample_rate * bytes_per_sample)
