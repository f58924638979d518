#!/usr/bin/env python
"""Open-loop streaming benchmark (the shape of BASELINE.json configs[2]: continuous batching under Poisson arrivals
on 1/2/4/8 replicas), with the metric definitions of the reference's own client (benchmark/goodput.py):

  * arrivals: inter-arrival ~ Gamma(k = burstiness = 1, theta = 1/rate), seed 42          (goodput.py:354-363)
  * TTFA: first audio chunk delivered - request arrival                                      (goodput.py:250-262)
  * streaming viability: share of chunks whose cumulative audio duration exceeds their arrival latency since the
    first chunk, and share of requests for which that holds for every chunk                  (goodput.py:186-215)
  * audio-sec/sec: total audio of completed requests / (last completion - first arrival)     (throughput.py:315-318)

    python benchmarks/openloop.py --rate 12 --duration 20                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        benchmarks/openloop.py --rate 96 --duration 20                        # 8 replicas, 96 req/s in total

One process per GPU.  Requests are pinned to replicas by ``vox_serve_b200.router.ReplicaRouter`` (round-robin = the
reference's rule, launch.py:471-474); every rank evaluates the same deterministic assignment, so the data path has no
cross-process traffic -- exactly like the reference's data-parallel mode, where the API server only forwards bytes.
What this run can show and the closed-loop bench.py cannot: queueing, the host-side cost of joins / leaves, TTFA under
load.  Model: Orpheus-3B with synthetic weights (the only adapter on the CUDA path), ``--tokens`` decode steps per
request (stop id masked), default sampling."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def arrivals(rate: float, duration: float, seed: int = 42, burstiness: float = 1.0):
    import numpy as np

    rng = np.random.default_rng(seed)
    t, out = 0.0, []
    while True:
        t += float(rng.gamma(burstiness, 1.0 / (burstiness * rate)))
        if t >= duration:
            return out
        out.append(t)


def viability(arr, dur):
    """(per-chunk %, all-chunks-ok) of one request, goodput.py:186-215."""
    if len(arr) < 2:
        return None
    ok = sum(1 for i in range(1, len(arr)) if sum(dur[:i]) > arr[i] - arr[0])
    return 100.0 * ok / (len(arr) - 1), ok == len(arr) - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rate", type=float, default=8.0, help="requests per second over ALL replicas")
    ap.add_argument("--duration", type=float, default=20.0, help="arrival window in seconds")
    ap.add_argument("--tokens", type=int, default=700, help="decode steps per request (100 SNAC frames = 8.53 s audio)")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--policy", default="round_robin", choices=["round_robin", "least_outstanding"])
    ap.add_argument("--sync", action="store_true", help="Scheduler._step instead of _step_async")
    ap.add_argument("--gpus", type=int, default=None, help="(informational; the world size comes from torchrun)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # its one-line banner goes to stdout regardless
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vox_serve_b200.model.orpheus import OrpheusModel
    from vox_serve_b200.requests import Request
    from vox_serve_b200.router import ReplicaRouter
    from vox_serve_b200.scheduler import Scheduler
    from vox_serve_b200.worker import ModelWorker

    prompt_tokens = 128
    max_tokens = prompt_tokens + 5 + args.tokens
    torch.manual_seed(1234 + rank)
    model = OrpheusModel(f"orpheus-synthetic:{rank}", device=f"cuda:{local}", mask_stop_token=True, max_tokens=max_tokens)
    pages = max(1024, args.batch * ((max_tokens + 127) // 128 + 1))
    worker = ModelWorker("orpheus-synthetic", max_batch_size=args.batch, max_num_pages=pages, page_size=128, model=model,
                         max_prefill_tokens=1024)
    worker.capture_decode_graphs()
    # ---- the whole arrival stream, identical on every rank; this rank keeps what the router pins to it ----
    times = arrivals(args.rate, args.duration)
    router = ReplicaRouter(world, args.policy if args.policy == "round_robin" else "round_robin")
    g = torch.Generator().manual_seed(42)
    mine = []
    for i, t in enumerate(times):
        ids = torch.randint(0, 128000, (prompt_tokens,), generator=g).tolist()
        if router.assign(i) == rank:
            mine.append((t, f"q{i}", ids))
    chunks = {}

    def on_audio(req, chunk, now):
        c = chunks.setdefault(req.request_id, ([], []))
        c[0].append(now)
        c[1].append(len(chunk) / 48000.0)

    sched = Scheduler(worker, on_audio=on_audio)
    # warm-up: one short untimed request exercises prefill + decode + vocoder once
    w = Request(request_id="warm", prompt=mine[0][2] if mine else [1] * prompt_tokens, model_kwargs={"voice": None})
    sched.submit(w)
    n = 0
    while not sched.audio["warm"] and n < 200:
        sched._step()
        n += 1
    worker.free_kv_cache(w)
    sched.active_requests, sched.pending = [], type(sched.pending)()
    chunks.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    arrive, finish = {}, {}
    state, nxt, steps = (None, [], []), 0, 0
    done_seen = 0
    while nxt < len(mine) or sched.has_work() or state[0] is not None:
        now = time.perf_counter() - t0
        while nxt < len(mine) and mine[nxt][0] <= now:
            t, rid, ids = mine[nxt]
            sched.submit(Request(request_id=rid, prompt=ids, is_streaming=True, model_kwargs={"voice": None}))
            arrive[rid] = t0 + t                     # the client's clock starts at the scheduled arrival
            nxt += 1
        if not sched.has_work() and state[0] is None:
            time.sleep(min(0.0005, max(0.0, mine[nxt][0] - now))) if nxt < len(mine) else None
            continue
        if args.sync:
            sched._step()
        else:
            state = sched._step_async(*state)
        steps += 1
        while done_seen < len(sched.finished):
            finish[sched.finished[done_seen].request_id] = time.perf_counter()
            done_seen += 1
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    rows = []
    for _, rid, _ in mine:
        arr, dur = chunks.get(rid, ([], []))
        v = viability(arr, dur)
        rows.append({"id": rid, "ttfa": (arr[0] - arrive[rid]) if arr else None, "audio_s": sum(dur), "chunks": len(arr),
                     "via": v[0] if v else None, "via_all": v[1] if v else None, "arrive": arrive[rid] - t0,
                     "finish": finish.get(rid, t_end) - t0})
    allrows = [rows]
    if world > 1:
        allrows = [None] * world
        dist.all_gather_object(allrows, rows)
    if rank == 0:
        flat = [r for rr in allrows for r in rr]
        ok = [r for r in flat if r["ttfa"] is not None]
        ttfa = sorted(r["ttfa"] * 1e3 for r in ok)
        span = max(r["finish"] for r in ok) - min(r["arrive"] for r in ok) if ok else float("nan")
        via = [r["via"] for r in ok if r["via"] is not None]

        def pct(p):
            return ttfa[min(len(ttfa) - 1, int(p * len(ttfa)))] if ttfa else None

        print(json.dumps({
            "metric": "open-loop audio-sec/sec, TTFA, streaming viability", "n_gpus": world, "rate_req_s": args.rate,
            "arrival_window_s": args.duration, "requests": len(flat), "completed": len(ok),
            "audio_sec_per_sec": sum(r["audio_s"] for r in ok) / span if ok else None,
            "offered_audio_sec_per_sec": args.rate * args.tokens / 7 * 2048 / 24000,
            "ttfa_ms": {"p50": pct(0.5), "p90": pct(0.9), "p99": pct(0.99), "max": ttfa[-1] if ttfa else None},
            "streaming_viability_pct": {"per_chunk_mean": sum(via) / len(via) if via else None,
                                        "all_chunks": 100.0 * sum(1 for r in ok if r["via_all"]) / max(1, len(via))},
            "scheduler": "sync _step" if args.sync else "async _step_async", "policy": "round_robin",
            "per_request": {"tokens": args.tokens, "prompt_tokens": prompt_tokens + 5, "audio_s": args.tokens / 7 * 2048 / 24000},
            "model": "Orpheus-3B synthetic weights, batch <= %d per replica" % args.batch,
            "rank0_steps": steps, "wall_s": t_end - t0}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
