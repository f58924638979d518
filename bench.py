#!/usr/bin/env python
"""Headline benchmark: Orpheus-3B streaming TTS, batch 32 per GPU, SNAC codec path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one scheduler iteration in steady state: one LM decode step for the 32 running requests (one new
token each) plus, every 7th step, the SNAC decode + PCM16 of the newest 28-token window of every request.
Metric: audio-sec/sec = seconds of audio emitted / seconds elapsed (benchmark/throughput.py:315-318 of the
reference); 7 tokens = 2048 samples = 85.33 ms.

The timed steps are taken from the STEADY STATE of the continuous batch the config describes (32 requests x 700 decode
steps each): the requests are staggered in groups of four, 84 tokens apart, so the batch holds the mix of sequence
lengths a server at batch 32 holds (kv ~190 ... ~780, mean ~490) instead of 32 requests that all just finished prefill.

Two figures per run, same state, same K steps:
  value : device-resident loop (ModelWorker.run_lm_decode_resident) -- inputs in HBM, no host work between
          CUDA-graph replays, PCM left in HBM; timed with CUDA events.
  e2e   : the reference-facing worker API driven by Scheduler._step -- per step a pinned-host -> device copy of
          the page table / positions / slots and a device -> host read of the sampled ids, plus the PCM chunks
          device -> host; timed with CUDA events around the whole region (host gaps included).
TTFA (time to first audio chunk, benchmark/goodput.py:250-262) is reported for a lone request, for a request joining
31 running streams, and for a burst of 32 -- the burst with the reference's one-prefill-per-step policy and, separately,
with prompts batched into shared prefill steps.
N > 1: one replica per GPU (request-parallel, no data-path collective), barrier + max over ranks.
--impl reference: the CPU oracle port of the reference's path (oracle/), all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEC_PER_FRAME = 2048 / 24000.0
TOKENS_PER_FRAME = 7
BATCH = 32
PROMPT_TOKENS = 128
PROMPT_ROWS = PROMPT_TOKENS + 5            # [128259] + ids + [128009, 128260, 128261, 128257]  (orpheus.py:356-358)
SETUP_TOKENS = 28 + BATCH                  # every request past its first vocoded window (one prefill per step)
STAGGER_GROUP, STAGGER_STEPS = 4, 84       # steady-state mix: group j of four requests is 84 j tokens further along


def steady_state_kv_lens():
    """kv length of every request of the steady-state batch when the timed region starts (both arms use it)."""
    return [PROMPT_ROWS + SETUP_TOKENS + STAGGER_STEPS * (i // STAGGER_GROUP) for i in range(BATCH)]

WORKLOAD = "Orpheus-3B streaming TTS batch=32 on 1xB200 (SNAC codec path)"


def workload_config(world: int) -> dict:
    """The `config` object of the JSON line: ONE definition for both arms (`--impl b200` and `--impl reference`)."""
    kv = steady_state_kv_lens()
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "prompt_tokens": PROMPT_ROWS,
            "state": "steady-state continuous batch: request groups staggered 84 tokens apart "
                     f"(kv {min(kv)}..{max(kv)}, mean {sum(kv) / BATCH:.1f}, when the timed region starts)",
            "sampling": "top_p 0.8 T 0.6 repetition_penalty 1.3 (orpheus.py:260-268), stop id masked",
            "page_size": 128, "parallelism": f"dp{world} replicas",
            "weights": "seeded N(0,0.02) bf16 at Orpheus-3B shapes; SNAC 24 kHz shapes, seeded",
            "l2": "inputs larger than L2: 6.6 GB of weights stream from HBM every step (126 MB L2)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------------
# CPU oracle leg (cpu_baseline of the main arm; the whole of --impl reference)
# --------------------------------------------------------------------------------------------------------------
def cpu_oracle_sample(n_lm_steps: int = 2, kv_len: int = None, threads=None, warmup: int = 1):
    """Times the oracle port (oracle/: the reference's arithmetic on torch CPU) on a bounded sample of the same
    workload: `n_lm_steps` batch-32 decode steps of Orpheus-3B at kv_len ~ the GPU run's starting length, plus
    one SNAC decode of 32 windows; audio-sec/sec from the steady-state mix (1 SNAC per 7 LM steps)."""
    import torch

    from oracle import lm_ops, orpheus as oorph, sampler as osampler, snac as osnac

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    if kv_len is None:          # the mean of the GPU arm's steady-state mix
        kv_len = int(round(sum(steady_state_kv_lens()) / BATCH))
    dims = oorph.OrpheusDims()
    t0 = time.perf_counter()
    blk = torch.randn(1 << 24, dtype=torch.float32).mul_(0.02).to(torch.bfloat16)   # 32 MB seeded block, tiled

    def mat(*shape, scale=1.0):
        n = 1
        for s in shape:
            n *= s
        out = torch.empty(n, dtype=torch.bfloat16)
        for o in range(0, n, blk.numel()):
            m = min(blk.numel(), n - o)
            out[o:o + m] = blk[:m]
        return out.view(*shape) * scale if scale != 1.0 else out.view(*shape)

    H, I = dims.hidden_size, dims.intermediate_size
    hq, hkv = dims.num_attention_heads * dims.head_dim, dims.num_key_value_heads * dims.head_dim
    w = {"model.embed_tokens.weight": mat(dims.vocab_size, H), "model.norm.weight": torch.ones(H, dtype=torch.bfloat16),
         "lm_head.weight": mat(dims.vocab_size, H)}
    for i in range(dims.num_hidden_layers):
        n = oorph.layer_names(i)
        w[n["ln1"]] = torch.ones(H, dtype=torch.bfloat16)
        w[n["ln2"]] = torch.ones(H, dtype=torch.bfloat16)
        w[n["q"]], w[n["k"]], w[n["v"]], w[n["o"]] = mat(hq, H), mat(hkv, H), mat(hkv, H), mat(H, hq)
        w[n["gate"]], w[n["up"]], w[n["down"]] = mat(I, H), mat(I, H), mat(H, I)
    page = 128
    warmup = max(1, warmup)
    n_pages_req = (kv_len + n_lm_steps + warmup + page) // page
    kv = mat(dims.num_hidden_layers, BATCH * n_pages_req, 2, page, dims.num_key_value_heads, dims.head_dim)
    setup_s = time.perf_counter() - t0
    cfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1)
    rep = torch.zeros(BATCH, 1, 1, dims.vocab_size, dtype=torch.bool)
    ids = torch.randint(128266, 156938, (BATCH,))
    gen = torch.Generator().manual_seed(0)
    lm_times = []
    with torch.no_grad():
        for step in range(n_lm_steps + warmup):     # the first `warmup` steps are not timed
            L = kv_len + step + 1
            npg = (L + page - 1) // page
            ip = [i * npg for i in range(BATCH + 1)]
            ix = [r * n_pages_req + j for r in range(BATCH) for j in range(npg)]
            last = [L - (npg - 1) * page] * BATCH
            wr = lm_ops.PagedWrapperCPU("decode", page)
            wr.plan(ip, ix, last)
            t1 = time.perf_counter()
            logits = oorph.lm_forward(w, dims, ids, torch.full((BATCH,), L, dtype=torch.int32), wr, kv)
            out = oorph.sampling_step(logits[:, None, :], cfg, rep, generator=gen)
            dt = time.perf_counter() - t1
            ids = out[:, 0]
            if step >= warmup:
                lm_times.append(dt)
        scfg = osnac.SnacConfig()
        ssd = osnac.synth_state_dict(scfg, seed=1)
        win = torch.randint(128266, 156938, (BATCH, 28, 1))
        codes = oorph.audio_codes_from_window(win, dims)
        noises = [torch.randn(s) for s in osnac.noise_shapes(scfg, BATCH, 16)]
        t1 = time.perf_counter()
        wav = osnac.decode(ssd, scfg, codes, noises)
        pcm = (wav[:, :, 2048:4096].numpy() * 32767).astype("int16")
        snac_s = time.perf_counter() - t1
    lm_s = sum(lm_times) / len(lm_times)
    per_step = lm_s + snac_s / TOKENS_PER_FRAME
    value = BATCH * SEC_PER_FRAME / TOKENS_PER_FRAME / per_step
    return {"value": value, "unit": "audio-sec/sec", "cores": threads, "kind": "port",
            "sample": f"{n_lm_steps} timed batch-32 Orpheus-3B decode steps (oracle/orpheus.py lm_forward + sampler, "
                      f"torch CPU bf16, kv_len {kv_len}) + 1 SNAC decode of 32 windows (oracle/snac.py, fp32); "
                      f"steady-state mix 7 LM steps : 1 SNAC; weights tiled from a seeded 32 MB block",
            "lm_step_s": lm_s, "snac_decode_s": snac_s, "setup_s": setup_s, "ms_per_step": per_step * 1e3,
            "pcm_bytes": int(pcm.nbytes)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a step = one batch-32 decode step of the same workload on the host cores (~0.5 s each): --steps / --warmup are
    # honoured exactly up to 60 steps in total (~30 s); beyond that the sample is cut and `steps` says what ran
    n_warm = max(1, min(args.warmup, 10))
    n_steps = max(1, min(args.steps, 60 - n_warm))
    t0 = time.perf_counter()
    r = cpu_oracle_sample(n_lm_steps=n_steps, warmup=n_warm)
    # ONE batch-32 replica on the box's host cores, whatever --gpus says: the host does not grow with the GPU count
    line = {"impl": "reference", "metric": "audio-sec/sec", "value": r["value"], "unit": "audio-sec/sec",
            "n_gpus": args.gpus, "steps": n_steps, "warmup": n_warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.gpus),
            "note": "CPU port of the reference path (the reference has no CPU path); one batch-32 replica on all host "
                    "cores -- NOT multiplied by n_gpus; every request at the mean kv length of the steady-state mix",
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "audio-sec/sec", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's version / debug banner goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # its one-line banner goes to stdout regardless
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vox_serve_b200 import ops
    from vox_serve_b200.model.orpheus import OrpheusModel
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler
    from vox_serve_b200.worker import ModelWorker

    K, W = args.steps, max(args.warmup, 3)
    hop = TOKENS_PER_FRAME
    n_groups = BATCH // STAGGER_GROUP
    joins = args.ttfa_joins
    # longest-lived request: steady-state mix + the longer of (timed steps, the TTFA-join measurement)
    max_tokens = PROMPT_ROWS + SETUP_TOKENS + STAGGER_STEPS * (n_groups - 1) + W + max(K, 32 * joins) + 96
    torch.manual_seed(1234 + rank)
    model = OrpheusModel(f"orpheus-synthetic:{rank}", device=f"cuda:{local}", mask_stop_token=True, max_tokens=max_tokens)
    pages = max(2048, BATCH * ((max_tokens + 127) // 128 + 1))
    worker = ModelWorker("orpheus-synthetic", max_batch_size=BATCH, max_num_pages=pages, page_size=128, model=model,
                         max_prefill_tokens=1024)
    t_cap0 = time.perf_counter()
    n_graphs = worker.capture_decode_graphs()         # start-up work, like the reference's graph initialisation
    capture_s = time.perf_counter() - t_cap0
    g = torch.Generator().manual_seed(42 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fresh_batch(tag, vocoder_batch_steps=1, stagger=True, one_prefill_per_step=True):
        sched = Scheduler(worker, vocoder_batch_steps=vocoder_batch_steps, one_prefill_per_step=one_prefill_per_step)
        t_sub = {}
        reqs = []
        for i in range(BATCH):
            ids = torch.randint(0, 128000, (PROMPT_TOKENS,), generator=g).tolist()
            r = Request(request_id=f"{tag}{i}", prompt=ids, model_kwargs={"voice": None})
            reqs.append(r)
            sched.submit(r)
        # steady state: every request prefetched and holding >= 28 tokens (first window already vocoded)
        n = 0
        while any(len(r.lm_output_audio_tokens) < 28 or not sched.audio[r.request_id] for r in reqs):
            sched._step()
            n += 1
            assert n < 400
        if stagger:
            # steady-state mix (untimed): group j of four requests runs 84 j more decode steps than group 0
            for j in range(1, BATCH // STAGGER_GROUP):
                worker.run_lm_decode_resident(reqs[j * STAGGER_GROUP:], STAGGER_STEPS, detokenize=False)
            for r in reqs:      # no vocoder backlog: the newest completed window has been delivered
                idx = ((len(r.lm_output_audio_tokens) - 28) // hop) * hop
                r.audio_decode_idx, r.next_audio_decode_idx = [idx], [idx]
        return sched, reqs, n

    def drain(sched, reqs):
        for r in reqs:
            worker.free_kv_cache(r)
        sched.active_requests = []

    clocks = ClockSampler(local)
    # ------------------------------------------------ e2e through the worker API ------------------------------
    # one untimed ramp first, like a server's warm-up request burst: first-use costs (allocator growth for every
    # vocoder batch size of the ramp, lazy kernel loading) otherwise land in the burst TTFA and make it jump between
    # 250 and 600 ms from run to run
    sched, reqs, _ = fresh_batch("w", stagger=False)
    drain(sched, reqs)
    torch.cuda.synchronize()
    t_setup0 = time.perf_counter()
    sched, reqs, n_setup = fresh_batch("b", stagger=False)
    ttfa = sorted((sched.first_audio_time[r.request_id] - sched.submit_time[r.request_id]) * 1e3 for r in reqs)
    drain(sched, reqs)
    # the same burst with the prompts batched into as few prefill steps as fit max_prefill_tokens (a scheduling POLICY the
    # reference leaves selectable, scheduler/base.py:283-286; not its default): one untimed pass first
    ttfa_bp = []
    for tag in ("pw", "p"):
        sched, reqs, _ = fresh_batch(tag, stagger=False, one_prefill_per_step=False)
        ttfa_bp = sorted((sched.first_audio_time[r.request_id] - sched.submit_time[r.request_id]) * 1e3 for r in reqs)
        drain(sched, reqs)
    t_setup0 += 0.0
    sched, reqs, _ = fresh_batch("e")
    setup_s = time.perf_counter() - t_setup0
    def timed_api_loop(sched, reqs, n_steps, async_mode):
        """n_steps scheduler iterations through the worker API; returns (ms, audio seconds, launches, h2d, d2h)."""
        audio_bytes = [0]

        def on_audio(req, chunk, now):
            audio_bytes[0] += len(chunk)

        state = (None, [], [])
        if async_mode:
            state = sched.run_async(W + 1, state)      # first call only selects; then W real steps
        else:
            for _ in range(W):
                sched._step()
        sched.on_audio = on_audio
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0, lc0 = worker.gpu_launches, ops.launch_count()
        wall0 = time.perf_counter()
        ev0.record()
        if async_mode:
            state = sched.run_async(n_steps, state)
        else:
            for _ in range(n_steps):
                sched._step()
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        ms = max(ev0.elapsed_time(ev1), wall * 1e3)
        sched.on_audio = None
        if async_mode and state[0] is not None:      # finish the last step's request-state update
            try:
                state[0].send(None)
            except StopIteration:
                pass
        # eager launches (detokenize) are counted by the ctypes layer, graph nodes by the worker
        launches = (worker.gpu_launches - launches0) + (ops.launch_count() - lc0)
        n_detok = audio_bytes[0] / (BATCH * 4096.0)
        h2d = worker.staging.n * 4 + n_detok * worker.win_host.numel() * 4 / n_steps
        d2h = BATCH * 8 + audio_bytes[0] / n_steps
        return ms, audio_bytes[0] / 48000.0, launches, h2d, d2h

    clocks.start()
    e2e_ms, e2e_audio_s, e2e_launches, h2d_step, d2h_step = timed_api_loop(sched, reqs, K, async_mode=True)
    mean_kv = sum(r.kv_token_len for r in reqs) / BATCH - K / 2
    drain(sched, reqs)
    # the same loop with the synchronous Scheduler._step (host bookkeeping serialised with the device)
    Ks = min(K, 140)
    sched, reqs, _ = fresh_batch("s")
    sync_ms, sync_audio_s, _, _, _ = timed_api_loop(sched, reqs, Ks, async_mode=False)
    drain(sched, reqs)
    # the async loop with the scheduler's optional vocoder batching (completed windows held until every 7th step so
    # the vocoder runs on the full batch; first chunks are never held): reported beside the default schedule
    sched, reqs, _ = fresh_batch("v", vocoder_batch_steps=hop)
    vb_ms, vb_audio_s, _, _, _ = timed_api_loop(sched, reqs, K, async_mode=True)
    drain(sched, reqs)

    # ------------------------------------------------ device-resident loop ------------------------------------
    sched, reqs, _ = fresh_batch("r")
    worker.run_lm_decode_resident(reqs, W + (hop - W % hop) % hop, detokenize=True)
    barrier()
    timing = {"start": torch.cuda.Event(enable_timing=True), "end": torch.cuda.Event(enable_timing=True)}
    launches0 = worker.gpu_launches
    worker.run_lm_decode_resident(reqs, K, detokenize=True, timing=timing, tail_window=True)
    torch.cuda.synchronize()
    res_ms = timing["start"].elapsed_time(timing["end"])
    res_launches = worker.gpu_launches - launches0
    # K steps = K / hop frames per request; a K that is not a multiple of the hop still pays ceil(K / hop) vocoder
    # passes inside the timed region (tail_window), so the figure can only err low
    res_audio_s = (K / hop) * BATCH * SEC_PER_FRAME
    clk = clocks.stop()
    if args.profile_steps > 0:       # ncu --profile-from-start off: only these replays are captured
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        worker.run_lm_decode_resident(reqs, args.profile_steps, detokenize=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ------------------------------------------------ kernel-isolated roofline passes -------------------------
    roof = kernel_rooflines(worker, model, reqs, torch, ops)
    drain(sched, reqs)

    # ------------------------------------------------ TTFA of a lone request on the warm server ----------------
    ttfa_single = []
    for i in range(5):
        s1 = Scheduler(worker)
        ids = torch.randint(0, 128000, (PROMPT_TOKENS,), generator=g).tolist()
        r1 = Request(request_id=f"t{i}", prompt=ids, model_kwargs={"voice": None})
        torch.cuda.synchronize()
        s1.submit(r1)
        n = 0
        while not s1.audio[r1.request_id]:
            s1._step()
            n += 1
            assert n < 200
        ttfa_single.append((s1.first_audio_time[r1.request_id] - s1.submit_time[r1.request_id]) * 1e3)
        worker.free_kv_cache(r1)
        s1.active_requests = []
    ttfa_single.sort()

    # ------------------------------------------------ steady-state TTFA: one request joins 31 running streams ------
    # ("define TTFA as steady-state (one request joining 31 running streams)"; goodput.py:250-262: first
    # audio chunk arrival - request start.)  The joining prefill rides in the same step as the 31 decodes.
    ttfa_join = []
    if joins > 0:
        sj, rj, _ = fresh_batch("j")
        victim = rj.pop()                       # 31 keep running
        worker.free_kv_cache(victim)
        sj.active_requests = [r for r in sj.active_requests if r is not victim]
        for _ in range(W):
            sj._step()
        for i in range(joins):
            ids = torch.randint(0, 128000, (PROMPT_TOKENS,), generator=g).tolist()
            r1 = Request(request_id=f"join{i}", prompt=ids, model_kwargs={"voice": None})
            torch.cuda.synchronize()
            sj.submit(r1)
            n = 0
            while not sj.audio[r1.request_id]:
                sj._step()
                n += 1
                assert n < 200
            ttfa_join.append((sj.first_audio_time[r1.request_id] - sj.submit_time[r1.request_id]) * 1e3)
            worker.free_kv_cache(r1)
            sj.active_requests = [r for r in sj.active_requests if r is not r1]
        drain(sj, rj)
        ttfa_join.sort()

    # ------------------------------------------------ reduce over ranks ----------------------------------------
    from vox_serve_b200.router import reduce_job_metrics

    (res_ms, e2e_ms), (res_audio_s, e2e_audio_s, res_launches, e2e_launches) = reduce_job_metrics(
        [res_ms, e2e_ms], [res_audio_s, e2e_audio_s, float(res_launches), float(e2e_launches)], device="cuda")
    if rank == 0:
        peak, peak_src = peaks()
        line = {
            "metric": "audio-sec/sec", "value": res_audio_s / (res_ms / 1e3), "unit": "audio-sec/sec", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": res_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world), "mean_kv_len_timed": round(mean_kv, 1),
            "e2e": {"value": e2e_audio_s / (e2e_ms / 1e3), "unit": "audio-sec/sec", "h2d_bytes_per_step": int(h2d_step),
                    "d2h_bytes_per_step": int(d2h_step), "ms_per_step": e2e_ms / K,
                    "api": "Scheduler._step_async (scheduler/base.py:168-215 ordering) -> ModelWorker."
                           "prepare_lm_inputs/run_detokenize/run_lm_decode; host bookkeeping one step behind the device",
                    "sync_scheduler": {"value": sync_audio_s / (sync_ms / 1e3) * world, "ms_per_step": sync_ms / Ks,
                                       "steps": Ks, "api": "Scheduler._step (scheduler/base.py:135-166)"},
                    "vocoder_batched": {"value": vb_audio_s / (vb_ms / 1e3) * world, "ms_per_step": vb_ms / K,
                                        "policy": f"Scheduler(vocoder_batch_steps={hop}): later chunks of a request "
                                                  "wait up to 6 steps for a full vocoder batch (not the default)"}},
            "gpu_launches": int(res_launches), "gpu_launches_e2e": int(e2e_launches),
            "tokens_per_s": BATCH * world * K / (res_ms / 1e3),
            "ttfa_burst_ms": {"p50": ttfa[len(ttfa) // 2], "min": ttfa[0], "max": ttfa[-1],
                              "note": "32 requests submitted at once to a warm replica (one untimed burst first), one prefill per"
                                      " step (scheduler/base.py:283-284); decode graphs captured at start-up"},
            "ttfa_burst_batched_prefill_ms": {"p50": ttfa_bp[len(ttfa_bp) // 2], "min": ttfa_bp[0], "max": ttfa_bp[-1],
                                              "note": "the same burst with Scheduler(one_prefill_per_step=False): prompts share "
                                                      "prefill steps up to max_prefill_tokens rows (policy, not the reference's "
                                                      "default)"},
            "ttfa_single_ms": {"p50": ttfa_single[len(ttfa_single) // 2], "min": ttfa_single[0], "max": ttfa_single[-1],
                               "note": "one 133-token request on an otherwise idle, warm replica: prefill + the 28 decode "
                                       "steps the first SNAC window needs + vocoder + PCM copy"},
            "ttfa_join_ms": ({"p50": ttfa_join[len(ttfa_join) // 2], "p90": ttfa_join[int(len(ttfa_join) * 0.9)],
                              "min": ttfa_join[0], "max": ttfa_join[-1], "joins": len(ttfa_join),
                              "note": "one 133-token request joins 31 running streams (steady-state kv mix): its prefill "
                                      "rides with the 31 decodes, then 27 batch-32 decode steps, vocoder, PCM copy"}
                             if ttfa_join else None),
            "setup_s": setup_s, "graph_capture_s": capture_s, "decode_graphs": n_graphs, "clocks": clk,
        }
        line.update(roof(peak, peak_src))
        if world == 1 and not args.no_cpu:
            try:
                line["cpu_baseline"] = {k: v for k, v in cpu_oracle_sample(2).items()
                                        if k in ("value", "unit", "cores", "kind", "sample", "lm_step_s", "snac_decode_s")}
            except Exception as e:      # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"value": None, "unit": "audio-sec/sec", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def kernel_rooflines(worker, model, reqs, torch, ops):
    """Times, with CUDA events on the launching stream, (a) all projection GEMM launches of one decode step and
    (b) the 28 paged-attention launches of one step, each as a CUDA graph replayed back to back on the live
    buffers (weights of 28 different layers: 6.6 GB per replay, far beyond L2).  Average launch duration =
    replay time / launches; achieved = algorithmic bytes / time."""
    eng = model.engine_for(worker.kv_cache, worker.page_size)
    d, B = eng.dims, len(reqs)
    H, I, V = d.hidden_size, d.intermediate_size, d.vocab_size
    hq, hkv, D = d.num_attention_heads, d.num_key_value_heads, d.head_dim
    kv_lens = [r.kv_token_len for r in reqs]

    def graph_of(fn):
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(gph, stream=s):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        return gph

    def time_graph(gph, reps=5):
        for _ in range(2):
            gph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n_gemm = [0]

    def gemms():
        n0 = ops.launch_count()
        eng.gemm_pass(B)
        n_gemm[0] = ops.launch_count() - n0

    def attns():
        for i in range(d.num_hidden_layers):
            eng.attention_only(i, B, worker.decode_wrapper.plan_rows)

    def gate_ups():
        for L in eng.w.layers:
            eng.gate_up_only(L, B)

    g1 = graph_of(gemms)
    gemm_ms = time_graph(g1)
    g3 = graph_of(gate_ups)
    gu_ms = time_graph(g3)
    gu_bytes = d.num_hidden_layers * (2 * I * H * 2 + B * 2 * (H + I))
    g2 = graph_of(attns)
    attn_ms = time_graph(g2)
    w_bytes = eng.w.streamed_bytes_per_step()
    act_bytes = d.num_hidden_layers * B * 2 * (H + (hq + 2 * hkv) * D + hq * D + H + H + I + I + H) + B * 2 * (H + V)
    gemm_bytes = w_bytes + act_bytes
    attn_bytes = d.num_hidden_layers * (sum(kv_lens) * 2 * hkv * D * 2 + 2 * B * hq * D * 2)

    def fill(peak, peak_src):
        ga = gemm_bytes / (gemm_ms * 1e-3) / 1e9
        aa = attn_bytes / (attn_ms * 1e-3) / 1e9
        gua = gu_bytes / (gu_ms * 1e-3) / 1e9
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ncu = json.load(f)
        except Exception:
            ncu = {}
        gu_traffic = ncu.get("gemm_gate_up", {}).get("dram_bytes_per_launch")
        mean_kv = sum(kv_lens) / len(kv_lens)
        caps = ncu.get("paged_attn") or []
        cap = min(caps, key=lambda c: abs(c["kv_len"] - mean_kv)) if caps else {}
        return {
            "roofline_gate_up": {"kernel": "gemm_bf16_kernel, gate/up projection launches only (28 per step, the largest "
                                           "projection: 100.7 MB of weights each)", "bound": "hbm", "achieved": gua,
                                 "peak": peak, "unit": "GB/s", "frac": gua / peak, "traffic": gu_traffic,
                                 "traffic_source": ncu.get("gemm_gate_up", {}).get("source"),
                                 "avg_launch_us": gu_ms * 1e3 / d.num_hidden_layers,
                                 "algorithmic_bytes_per_launch": int(gu_bytes / d.num_hidden_layers)},
            "roofline": {"kernel": "gemm_bf16_kernel (all projection launches of one decode step)", "bound": "hbm",
                         "achieved": ga, "peak": peak, "unit": "GB/s", "frac": ga / peak,
                         "traffic": ncu.get("gemm_all", {}).get("dram_bytes_per_launch"),
                         "traffic_per_step": ncu.get("gemm_all", {}).get("dram_bytes_per_step"),
                         "traffic_source": ncu.get("gemm_all", {}).get("source"),
                         "peak_source": peak_src, "launches_per_step": n_gemm[0],
                         "avg_launch_us": gemm_ms * 1e3 / max(1, n_gemm[0]),
                         "algorithmic_bytes_per_step": int(gemm_bytes)},
            "roofline_attention": {"kernel": "paged_attn_kernel (28 launches of one decode step)", "bound": "hbm",
                                   "achieved": aa, "peak": peak, "unit": "GB/s", "frac": aa / peak,
                                   "traffic": cap.get("dram_bytes_per_launch"),
                                   "traffic_source": (f"{cap.get('source')}: the capture nearest in kv length (all 32 rows at "
                                                      f"kv {cap.get('kv_len')}, {cap.get('algorithmic_bytes_per_launch')} "
                                                      "algorithmic bytes per launch)") if cap else None,
                                   "avg_launch_us": attn_ms * 1e3 / d.num_hidden_layers,
                                   "mean_kv_len": mean_kv,
                                   "algorithmic_bytes_per_launch": int(attn_bytes / d.num_hidden_layers),
                                   "algorithmic_bytes_per_step": int(attn_bytes)},
        }

    return fill


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=140,
                    help="timed steps; the batch is not refilled inside the timed region, so long runs drift towards "
                         "longer contexts than the steady-state mix (700 steps: mean kv 830 instead of 490-560)")
    ap.add_argument("--warmup", type=int, default=7)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ttfa-joins", type=int, default=20, help="joins measured for ttfa_join_ms (0 = skip)")
    ap.add_argument("--profile-steps", type=int, default=0,
                    help="after the timed regions run this many resident steps inside cudaProfilerStart/Stop (ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
