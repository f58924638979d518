"""The KV hand-off protocol (vox_serve_b200/kv_handoff.py) over a world_size-2 gloo process group on CPU: message
order, page re-mapping on the receiver, slot-resident state, capacity errors that leave the channel in sync, and the
broadcast fan-out.  The device half (vb_copy_pages) is replaced by torch indexing injected by the test -- the product
defaults refuse CPU tensors; the GPU tests cover them."""
import os
import queue
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vox_serve_b200._lib import VoxB200Error
from vox_serve_b200.kv_handoff import DistTransport, KVHandoff, LoopbackTransport
from vox_serve_b200.requests import Request

L, PAGES, PAGE, HKV, D, V, B = 2, 8, 4, 1, 8, 50, 3


class _Worker:
    """The attribute surface of ModelWorker that the hand-off touches, on CPU tensors."""
    has_depth_transformer = False

    def __init__(self, seed: int, rep: bool = True):
        g = torch.Generator().manual_seed(seed)
        self.kv_cache = torch.randn(L, PAGES, 2, PAGE, HKV, D, generator=g).to(torch.bfloat16)
        self.page_size, self.max_num_pages, self.max_batch_size = PAGE, PAGES, B
        self.empty_pages = queue.Queue()
        for i in range(PAGES):
            self.empty_pages.put(i)
        self.free_slots, self.slot_of = list(range(B - 1, -1, -1)), {}
        self.next_input = torch.zeros(B, dtype=torch.int32)
        self.n_out = torch.zeros(B, dtype=torch.int32)
        self.history = torch.zeros(B, 64, dtype=torch.int32)
        self.rep_cache = torch.zeros(B, 1, 1, V, dtype=torch.bool) if rep else None

    def _acquire_slot(self, req):
        if req.request_id not in self.slot_of:
            self.slot_of[req.request_id] = self.free_slots.pop()
        return self.slot_of[req.request_id]

    def free_kv_cache(self, req):
        for p in req.kv_pages or []:
            self.empty_pages.put(p)
        req.kv_pages, req.kv_token_len, req.kv_last_page_len = [], 0, 0
        s = self.slot_of.pop(req.request_id, None)
        if s is not None:
            self.free_slots.append(s)


def _pack(cache, ids):
    return cache[:, ids.long()].contiguous()


def _unpack(cache, ids, staging):
    cache[:, ids.long()] = staging


def _prefilled(worker, rid="req-é1", pages=(5, 2), n_out=6):
    """A request as it looks after prefill + n_out decode steps on `worker`."""
    req = Request(request_id=rid)
    slot = worker._acquire_slot(req)
    taken = [worker.empty_pages.get_nowait() for _ in range(PAGES)]
    for p in taken:
        if p not in pages:
            worker.empty_pages.put(p)
    req.kv_pages = list(pages)
    req.kv_token_len, req.kv_last_page_len = (len(pages) - 1) * PAGE + 3, 3
    req.next_position_id, req.input_length, req.done_lm_prefill = req.kv_token_len + 1, 2, True
    worker.next_input[slot] = 41
    worker.n_out[slot] = n_out
    worker.history[slot, :n_out] = torch.arange(100, 100 + n_out, dtype=torch.int32)
    if worker.rep_cache is not None:
        worker.rep_cache[slot, 0, 0, [3, 7, 41]] = True
    req.next_audio_decode_idx = [7]
    return req


def _check_arrived(req, worker, src_cache, src_pages, n_out=6):
    assert req.request_id == "req-é1" and req.done_lm_prefill and not req.done_lm_generation
    assert len(req.kv_pages) == len(src_pages)
    assert torch.equal(worker.kv_cache[:, req.kv_pages], src_cache[:, list(src_pages)])
    slot = worker.slot_of[req.request_id]
    assert int(worker.next_input[slot]) == 41 and int(worker.n_out[slot]) == n_out
    assert worker.history[slot, :n_out].tolist() == list(range(100, 100 + n_out))
    assert [int(t[0, 0]) for t in req.lm_output_tokens] == list(range(100, 100 + n_out))
    assert req.lm_output_tokens[0].shape == (1, 1) and req.lm_output_tokens[0].dtype == torch.int64
    assert req.kv_last_page_len == 3 and req.kv_token_len == (len(src_pages) - 1) * PAGE + 3
    assert req.next_position_id == req.kv_token_len + 1 and req.input_length == 2
    assert req.next_audio_decode_idx == [7] and int(req.input_tokens[0, 0]) == 41
    if worker.rep_cache is not None:
        assert torch.nonzero(worker.rep_cache[slot, 0, 0]).flatten().tolist() == [3, 7, 41]


def test_loopback_roundtrip_and_release():
    a, b = _Worker(1), _Worker(2)
    src = a.kv_cache.clone()
    wire = LoopbackTransport()
    ha, hb = KVHandoff(a, wire, _pack, _unpack), KVHandoff(b, wire, _pack, _unpack)
    req = _prefilled(a)
    b.empty_pages.get_nowait()                  # receiver's next free pages are 1, 2: not the sender's 5, 2
    n = ha.send_request(req, dst=1)
    assert a.empty_pages.qsize() == PAGES and len(a.free_slots) == B and req.kv_pages == []
    got = hb.recv_request(src=0)
    assert got.kv_pages == [1, 2]
    _check_arrived(got, b, src, (5, 2))
    assert n == ha.bytes_sent and hb.bytes_received == n - len("req-é1".encode())


def test_capacity_error_keeps_the_channel_in_sync():
    a, b = _Worker(1), _Worker(2)
    wire = LoopbackTransport()
    ha, hb = KVHandoff(a, wire, _pack, _unpack), KVHandoff(b, wire, _pack, _unpack)
    req = _prefilled(a)
    held = [b.empty_pages.get_nowait() for _ in range(PAGES - 1)]      # one page free, two needed
    ha.send_request(req, dst=1, release=False)
    with pytest.raises(VoxB200Error, match="pages needed"):
        hb.recv_request(src=0)
    assert not wire.fifo and b.empty_pages.qsize() == 1 and len(b.free_slots) == B
    for p in held:
        b.empty_pages.put(p)
    ha.send_request(req, dst=1)                 # the sender still held it: the retry goes through
    _check_arrived(hb.recv_request(src=0), b, a.kv_cache, (5, 2))
    with pytest.raises(VoxB200Error, match="no prefilled KV"):
        ha.send_request(req, dst=1)             # released: nothing left to send


def test_product_defaults_refuse_cpu_tensors():
    a = _Worker(1)
    h = KVHandoff(a, LoopbackTransport())
    with pytest.raises(Exception):
        h.send_request(_prefilled(a), dst=1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = _Worker(seed=10 + rank, rep=True)
        h = KVHandoff(w, DistTransport(), _pack, _unpack)
        src_cache = _Worker(seed=10).kv_cache      # what rank 0 holds (same seed)
        # ---- point to point: 0 -> 1 ----
        if rank == 0:
            req = _prefilled(w)
            h.send_request(req, dst=1)
            ok = w.empty_pages.qsize() == PAGES
        else:
            w.empty_pages.get_nowait()
            got = h.recv_request(src=0)
            _check_arrived(got, w, src_cache, (5, 2))
            ok = got.kv_pages == [1, 2]
            w.free_kv_cache(got)
        # ---- fan-out: rank 0's prompt KV to every replica ----
        if rank == 0:
            req = _prefilled(w, pages=(6, 0, 3), n_out=1)
            kept = h.broadcast_request(req, src=0, rank=0)
            ok = ok and kept is req and req.kv_pages == [6, 0, 3]
        else:
            got = h.broadcast_request(None, src=0, rank=rank)
            _check_arrived(got, w, src_cache, (6, 0, 3), n_out=1)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if rank == 0:
            out.put(flags)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_handoff_and_broadcast():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    flags = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flags == [True, True]
