"""Prefill-KV hand-off between replicas: a request that was prefilled (or partly decoded) on one GPU continues on
another (north_star: "NCCL over NVLink only for the optional STS prefill-KV broadcast"; SURVEY.md §8e: point-to-point
send / recv -- or a broadcast when one prompt's KV fans out to several replicas -- of ``[L, n_pages, 2, page, Hkv, D]``
slices).  The reference has no counterpart: it pins a request to one replica for its lifetime
(``vox_serve/launch.py:471-474``) and its disaggregation mode moves only token ids between an LM and a detokenizer
device (``scheduler/disaggregation.py``).  This is the one place the data path uses a collective, and it is off the
steady-state loop: a replica that is not handing a request over never calls into it.

What travels (one message sequence per request, every tensor on the transport's device):

1. header, int64 [HEADER_LEN]: page count, kv length, last-page fill, next position id, prompt length, tokens
   generated so far, the token the next decode step feeds, repetition-cache bytes, request-id bytes, vocoder progress;
2. the request id (uint8);
3. the K/V pages of every layer as ONE contiguous tensor ``[L, n_pages, 2, page, Hkv, D]`` -- gathered from the
   sender's paged cache by ``vb_copy_pages`` and scattered into freshly allocated pages of the receiver's cache by the
   same kernel (page ids differ between replicas; only the receiver's page table knows the new ones);
4. the slot-resident decode state of ``ModelWorker``: the token history ring (int32 [n_out]) and the repetition-cache
   row (uint8), when present.

Transports: ``DistTransport`` (``torch.distributed`` send / recv / broadcast: NCCL over NVLink between the GPUs of a
box, gloo in the CPU tests of the protocol) and ``LoopbackTransport`` (two workers in one process).  The single-codebook
``ModelWorker`` only (Orpheus / GLM-style LMs -- the STS configuration); multi-codebook workers raise.
"""
from __future__ import annotations

from collections import deque
from typing import Callable, List, Optional

import torch

from ._lib import VoxB200Error
from .requests import Request

HEADER_LEN = 12
(H_PAGES, H_KV_LEN, H_LAST, H_NEXT_POS, H_INPUT_LEN, H_N_OUT, H_NEXT_INPUT, H_REP_BYTES, H_ID_BYTES, H_AUDIO_IDX,
 H_DONE_LM, H_MAGIC) = range(HEADER_LEN)
MAGIC = 0x4B564832       # "KVH2"


class DistTransport:
    """torch.distributed point-to-point / broadcast on the default (or a given) process group."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise VoxB200Error("DistTransport needs an initialised torch.distributed process group")
        self.dist, self.group = dist, group

    def send(self, t: torch.Tensor, dst: int) -> None:
        self.dist.send(t, dst, group=self.group)

    def recv(self, t: torch.Tensor, src: int) -> None:
        self.dist.recv(t, src, group=self.group)

    def broadcast(self, t: torch.Tensor, src: int) -> None:
        self.dist.broadcast(t, src, group=self.group)


class LoopbackTransport:
    """Two workers of ONE process on one device: the "wire" is a FIFO of device tensors."""

    def __init__(self):
        self.fifo = deque()

    def send(self, t: torch.Tensor, dst: int) -> None:
        self.fifo.append(t.clone())

    def recv(self, t: torch.Tensor, src: int) -> None:
        t.copy_(self.fifo.popleft())

    def broadcast(self, t: torch.Tensor, src: int) -> None:      # a group of one
        pass


def _default_pack(kv_cache: torch.Tensor, page_ids: torch.Tensor) -> torch.Tensor:
    from . import ops

    return ops.copy_pages(kv_cache, page_ids)


def _default_unpack(kv_cache: torch.Tensor, page_ids: torch.Tensor, staging: torch.Tensor) -> None:
    from . import ops

    ops.copy_pages(kv_cache, page_ids, staging, to_cache=True)


class KVHandoff:
    """Moves requests between ``ModelWorker`` replicas.  ``pack`` / ``unpack`` default to the CUDA page gather /
    scatter (``ops.copy_pages``; they raise on anything but CUDA tensors); the protocol tests inject their own."""

    def __init__(self, worker, transport, pack: Optional[Callable] = None, unpack: Optional[Callable] = None):
        if getattr(worker, "has_depth_transformer", False):
            raise VoxB200Error("KV hand-off serves single-codebook LM workers only")
        self.worker, self.transport = worker, transport
        self.pack = pack or _default_pack
        self.unpack = unpack or _default_unpack
        self.bytes_sent = 0
        self.bytes_received = 0

    # ---- helpers --------------------------------------------------------------------------------------
    @property
    def device(self):
        return self.worker.kv_cache.device

    def _sync(self):
        """Device-wide: the worker's streams AND the transport's (NCCL runs its transfers on a stream of its own)."""
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def _message(self, req: Request) -> List[torch.Tensor]:
        """header, id bytes, pages, history, repetition row of a request this worker holds."""
        w, dev = self.worker, self.device
        if req.request_id not in w.slot_of or not req.done_lm_prefill or not req.kv_pages:
            raise VoxB200Error(f"request {req.request_id} holds no prefilled KV on this worker")
        self._sync()                      # every step that touched the request has finished
        slot = w.slot_of[req.request_id]
        n_out = int(w.n_out[slot])
        rid = req.request_id.encode("utf-8")
        rep = None
        if w.rep_cache is not None:
            rep = w.rep_cache[slot].reshape(-1).view(torch.uint8)
        hdr = torch.zeros(HEADER_LEN, dtype=torch.int64)
        hdr[H_PAGES], hdr[H_KV_LEN], hdr[H_LAST] = len(req.kv_pages), req.kv_token_len, req.kv_last_page_len
        hdr[H_NEXT_POS], hdr[H_INPUT_LEN], hdr[H_N_OUT] = req.next_position_id, req.input_length or 0, n_out
        hdr[H_NEXT_INPUT] = int(w.next_input[slot])
        hdr[H_REP_BYTES] = 0 if rep is None else rep.numel()
        hdr[H_ID_BYTES] = len(rid)
        hdr[H_AUDIO_IDX] = req.next_audio_decode_idx[-1] if req.next_audio_decode_idx else -1
        hdr[H_DONE_LM] = int(bool(req.done_lm_generation))
        hdr[H_MAGIC] = MAGIC
        pages = self.pack(w.kv_cache, torch.tensor(req.kv_pages, dtype=torch.int32, device=dev))
        msg = [hdr.to(dev), torch.tensor(list(rid), dtype=torch.uint8, device=dev), pages]
        if n_out > 0:
            msg.append(w.history[slot, :n_out].contiguous())
        if rep is not None:
            msg.append(rep.contiguous())
        return msg

    def _adopt(self, hdr: torch.Tensor, rid: str, pages: Optional[torch.Tensor], hist: Optional[torch.Tensor],
               rep: Optional[torch.Tensor], req: Optional[Request]) -> Request:
        w, dev = self.worker, self.device
        n_pages, n_out = int(hdr[H_PAGES]), int(hdr[H_N_OUT])
        req = req or Request(request_id=rid)
        slot = w._acquire_slot(req)
        req.kv_pages = [w.empty_pages.get_nowait() for _ in range(n_pages)]
        self.unpack(w.kv_cache, torch.tensor(req.kv_pages, dtype=torch.int32, device=dev), pages)
        req.kv_token_len, req.kv_last_page_len = int(hdr[H_KV_LEN]), int(hdr[H_LAST])
        req.next_position_id, req.input_length = int(hdr[H_NEXT_POS]), int(hdr[H_INPUT_LEN])
        req.done_lm_prefill = True
        req.done_lm_generation = bool(int(hdr[H_DONE_LM]))
        w.next_input[slot] = int(hdr[H_NEXT_INPUT])
        w.n_out[slot] = n_out
        if n_out > 0:
            w.history[slot, :n_out] = hist
        if rep is not None:
            w.rep_cache[slot].reshape(-1).view(torch.uint8).copy_(rep)
            req.repetition_cache = w.rep_cache[slot]
        # host mirrors of the generated tokens ([1, n_codebooks] int64 rows, as sampling_host appends them)
        host = hist.to("cpu", torch.int64).view(-1, 1, 1) if n_out > 0 else torch.zeros(0, 1, 1, dtype=torch.int64)
        req.lm_output_tokens = [host[i] for i in range(n_out)]
        req.lm_output_audio_tokens = [host[i] for i in range(n_out)]
        req.input_tokens = torch.full((1, 1), int(hdr[H_NEXT_INPUT]), dtype=torch.int64, device=dev)
        a = int(hdr[H_AUDIO_IDX])
        req.next_audio_decode_idx = [a] if a >= 0 else []
        req.audio_decode_idx = list(req.next_audio_decode_idx)
        return req

    def _receive(self, xfer: Callable[[torch.Tensor], None], req: Optional[Request]) -> Request:
        w, dev = self.worker, self.device
        hdr = torch.zeros(HEADER_LEN, dtype=torch.int64, device=dev)
        xfer(hdr)
        hdr = hdr.cpu()
        if int(hdr[H_MAGIC]) != MAGIC:
            raise VoxB200Error("KV hand-off: header out of sync")
        rid_t = torch.zeros(int(hdr[H_ID_BYTES]), dtype=torch.uint8, device=dev)
        xfer(rid_t)
        rid = bytes(rid_t.cpu().tolist()).decode("utf-8")
        n_pages, n_out, rep_bytes = int(hdr[H_PAGES]), int(hdr[H_N_OUT]), int(hdr[H_REP_BYTES])
        pages = torch.empty((w.kv_cache.shape[0], n_pages) + tuple(w.kv_cache.shape[2:]), dtype=w.kv_cache.dtype,
                            device=dev)
        xfer(pages)
        hist = None
        if n_out > 0:
            hist = torch.empty(n_out, dtype=torch.int32, device=dev)
            xfer(hist)
        rep = None
        if rep_bytes > 0:
            rep = torch.empty(rep_bytes, dtype=torch.uint8, device=dev)
            xfer(rep)
        self.bytes_received += pages.numel() * pages.element_size() + 4 * n_out + rep_bytes + 8 * HEADER_LEN
        # the whole message has been taken off the wire before anything can fail: the channel stays in sync
        if w.empty_pages.qsize() < n_pages:
            raise VoxB200Error(f"KV hand-off: {n_pages} pages needed, {w.empty_pages.qsize()} free")
        if not w.free_slots and rid not in w.slot_of:
            raise VoxB200Error("KV hand-off: no free batch slot on the receiving replica")
        if n_out > w.history.shape[1] or (rep is not None and
                                          (w.rep_cache is None or rep_bytes != w.rep_cache[0].numel())):
            raise VoxB200Error("KV hand-off: sender and receiver are configured differently")
        return self._adopt(hdr, rid, pages, hist, rep, req)

    # ---- point to point -------------------------------------------------------------------------------
    def send_request(self, req: Request, dst: int, release: bool = True) -> int:
        """Hand ``req`` to replica ``dst``; with ``release`` its pages and batch slot return to this worker's pools
        (the caller drops it from its scheduler).  Returns the bytes put on the wire."""
        msg = self._message(req)
        for t in msg:
            self.transport.send(t, dst)
        n = sum(t.numel() * t.element_size() for t in msg)
        self.bytes_sent += n
        # returns once the receiver has taken the message: no transfer is left pending behind the caller's next steps
        # (CUDA-graph captures, page re-use), and the staging tensors may be freed
        self._sync()
        if release:
            self.worker.free_kv_cache(req)
        return n

    def recv_request(self, src: int, req: Optional[Request] = None) -> Request:
        """Take a request over from replica ``src``: allocate pages + a batch slot, scatter the K/V, restore the
        slot-resident decode state.  ``req``: an existing Request object to fill (e.g. the one the API layer holds)."""
        return self._receive(lambda t: self.transport.recv(t, src), req)

    # ---- fan-out ----------------------------------------------------------------------------------------
    def broadcast_request(self, req: Optional[Request], src: int, rank: int) -> Optional[Request]:
        """One prompt's KV to every replica of the group (the STS case: the same audio-prompt prefix serves several
        streams).  On ``src`` pass the request (it keeps it); every other rank passes None and gets its own copy."""
        if rank == src:
            for t in self._message(req):
                self.transport.broadcast(t, src)
            self._sync()
            return req
        return self._receive(lambda t: self.transport.broadcast(t, src), None)
