"""Prefill-KV hand-off between replicas (vox_serve_b200/kv_handoff.py) on the GPU: the page gather / scatter kernel
against torch indexing, a request migrating between two workers of one process (loopback transport), and -- when the
box has two GPUs -- the same migration over NCCL send / recv.  A migrated request must produce exactly the tokens and
PCM bytes it produces when it stays where it was prefilled (greedy, confident-model weights)."""
import os
import socket

import pytest
import torch

from oracle import orpheus as oorph, snac as osnac

pytestmark = pytest.mark.gpu

PROMPT_LENS, N_TOKENS = (9, 16, 12), 40


def test_copy_pages_matches_indexing():
    from vox_serve_b200 import ops

    g = torch.Generator().manual_seed(0)
    cache = torch.randn(3, 20, 2, 16, 2, 64, generator=g).to(torch.bfloat16).cuda()
    ids = torch.tensor([5, 17, 0, 9], dtype=torch.int32, device="cuda")
    staging = ops.copy_pages(cache, ids)
    assert torch.equal(staging, cache[:, ids.long()])
    other = torch.zeros_like(cache)
    new_ids = torch.tensor([1, 2, 19, 7], dtype=torch.int32, device="cuda")
    ops.copy_pages(other, new_ids, staging, to_cache=True)
    assert torch.equal(other[:, new_ids.long()], cache[:, ids.long()])
    mask = torch.ones(20, dtype=torch.bool)
    mask[new_ids.cpu().long()] = False
    assert torch.count_nonzero(other[:, mask.cuda()]) == 0
    # Orpheus-sized pages (512 KiB per page-layer), more pages than one sweep of the grid covers
    cache = torch.randn(2, 12, 2, 128, 8, 128, generator=g).to(torch.bfloat16).cuda()
    ids = torch.tensor([11, 3, 4, 8, 0], dtype=torch.int32, device="cuda")
    assert torch.equal(ops.copy_pages(cache, ids), cache[:, ids.long()])


def _replica(seed=3):
    from tests.e2e_harness import build_models

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    dims.max_tokens = max(PROMPT_LENS) + N_TOKENS
    worker, _ = build_models(dims, osnac.SnacConfig.tiny(), seed, len(PROMPT_LENS), 16, 128, planted=2.0)
    # deterministic vocoder: NoiseBlock noise injected as zeros (the device generator would differ between runs)
    worker.model.audio_decoder.noise_source = lambda shapes: [torch.zeros(s, device="cuda") for s in shapes]
    return worker, dims


def _prompts(dims):
    g = torch.Generator().manual_seed(21)
    return [torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist() for n in PROMPT_LENS]


def _tokens(req):
    return [int(t[0, 0]) for t in req.lm_output_tokens]


def _baseline(worker, dims, tag):
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    sched = Scheduler(worker)
    reqs = [Request(request_id=f"{tag}{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(_prompts(dims))]
    for r in reqs:
        sched.submit(r)
    sched.run_until_done(max_steps=4000)
    return [_tokens(r) for r in reqs], [sched.audio[r.request_id] for r in reqs]


def test_request_migrates_between_workers_loopback():
    from vox_serve_b200.kv_handoff import KVHandoff, LoopbackTransport
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    wa, dims = _replica()
    wb, _ = _replica()
    base_tokens, base_audio = _baseline(wa, dims, "b")
    assert wa.empty_pages.qsize() == wa.max_num_pages

    wire = LoopbackTransport()
    ha, hb = KVHandoff(wa, wire), KVHandoff(wb, wire)
    sa, sb = Scheduler(wa), Scheduler(wb)
    reqs = [Request(request_id=f"m{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(_prompts(dims))]
    for r in reqs:
        sa.submit(r)
    for _ in range(12):                     # three prefill steps, then every request has decoded a few tokens
        sa._step()
    mover = sa.detach("m1")
    assert mover.done_lm_prefill and len(mover.lm_output_tokens) >= 8 and not mover.done_lm_generation
    pages_a = list(mover.kv_pages)
    held = [wb.empty_pages.get_nowait() for _ in range(5)]          # the receiver's free pages differ from the sender's
    sent = ha.send_request(mover, dst=1)
    assert sent > 0 and wa.empty_pages.qsize() == wa.max_num_pages - sum(len(r.kv_pages) for r in reqs if r is not mover)
    arrived = hb.recv_request(src=0)
    for p in held:
        wb.empty_pages.put(p)
    assert arrived.request_id == "m1" and arrived.kv_pages != pages_a and len(arrived.kv_pages) == len(pages_a)
    assert _tokens(arrived) == _tokens(mover)
    sb.adopt(arrived)
    sa.run_until_done(max_steps=4000)
    sb.run_until_done(max_steps=4000)
    torch.cuda.synchronize()
    assert _tokens(reqs[0]) == base_tokens[0] and _tokens(reqs[2]) == base_tokens[2]
    assert _tokens(arrived) == base_tokens[1], "a migrated request must continue exactly where it left off"
    assert sa.audio["m1"] + sb.audio["m1"] == base_audio[1], "PCM of the migrated stream differs"
    assert sa.audio["m0"] == base_audio[0] and sa.audio["m2"] == base_audio[2]
    assert arrived.finish_reason == "max_tokens_reached"
    assert wb.empty_pages.qsize() == wb.max_num_pages and len(wb.free_slots) == wb.max_batch_size
    assert wa.empty_pages.qsize() == wa.max_num_pages and len(wa.free_slots) == wa.max_batch_size
    assert hb.bytes_received == ha.bytes_sent - len("m1")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_rank(rank, port, out):
    try:
        _nccl_rank_body(rank, port, out)
    except BaseException:            # the parent must hear about it: it polls the queue, not the exit codes
        import traceback

        out.put(("error", rank, traceback.format_exc()))
        raise


def _nccl_rank_body(rank, port, out):
    import torch.distributed as dist

    from vox_serve_b200.kv_handoff import DistTransport, KVHandoff
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=2, device_id=torch.device("cuda", rank))
    try:
        worker, dims = _replica()
        h = KVHandoff(worker, DistTransport())
        sched = Scheduler(worker)
        if rank == 0:
            reqs = [Request(request_id=f"m{i}", prompt=p, model_kwargs={"voice": None})
                    for i, p in enumerate(_prompts(dims))]
            for r in reqs:
                sched.submit(r)
            for _ in range(12):
                sched._step()
            mover = sched.detach("m1")
            n = h.send_request(mover, dst=1)          # returns when rank 1 has taken the message
            sched.run_until_done(max_steps=4000)
            out.put(("sent", n, sched.audio["m1"]))
        else:
            arrived = h.recv_request(src=0)
            sched.adopt(arrived)
            sched.run_until_done(max_steps=4000)
            torch.cuda.synchronize()
            tail = sched.audio["m1"]
            base_tokens, base_audio = _baseline(worker, dims, "b")
            out.put(("recv", _tokens(arrived) == base_tokens[1], tail, base_audio[1]))
    finally:
        dist.destroy_process_group()


def test_request_migrates_between_gpus_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import queue as pyqueue
    import time

    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_rank, args=(r, port, out), daemon=True) for r in range(2)]
    for p in procs:
        p.start()
    got, deadline = {}, time.time() + 150
    try:
        while len(got) < 2:
            try:
                m = out.get(timeout=2)
            except pyqueue.Empty:
                assert time.time() < deadline, "hand-off over NCCL did not finish in 150 s"
                assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a rank died without reporting"
                continue
            assert m[0] != "error", f"rank {m[1]} failed:\n{m[2]}"
            got[m[0]] = m[1:]
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()
    n_sent, head_audio = got["sent"]
    tokens_equal, tail_audio, base_audio = got["recv"]
    assert n_sent > 0 and tokens_equal
    assert head_audio + tail_audio == base_audio
