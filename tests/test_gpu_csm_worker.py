"""CSM served end to end through the worker API (``prepare_lm_inputs`` -> ``run_detokenize``
-> ``run_lm_prefill`` / ``run_lm_decode`` driven by the scheduler loop, vox_serve/worker/base.py:210-681 with the depth
path of :430-452, 510-614): staggered requests with ragged prompts, mixed prefill + decode steps, graph-replayed decode
frames, Mimi chunks every 10 frames, stop frames and the max_tokens guard.

The CPU oracle replays every request on its own, teacher-forced with the GPU's frames: every codebook of every frame must
be the oracle's argmax unless the oracle's own top-2 margin is a bf16 near-tie; every audio chunk must be the oracle's
Mimi decode of the same frames (int16, +-2 LSB for fp32 summation order)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import csm as ocsm, lm_ops, mimi as omimi

pytestmark = pytest.mark.gpu
TOL = 2e-2             # near-tie: top-2 margin below 2 * TOL * max|logit| (bf16 pipeline, same rule as test_gpu_csm.py)


def _build(seed, max_bs, stop_boost=None, max_tokens=None, pages=64, page=16):
    from vox_serve_b200.depth_engine import CsmDims
    from vox_serve_b200.model.csm import CSMModel
    from vox_serve_b200.sampling import SamplingConfig
    from vox_serve_b200.tokenizer.mimi import MimiConfig
    from vox_serve_b200.worker import CudaGraphWorker, DepthModelWorker

    odims = ocsm.CsmDims.tiny()
    weights = ocsm.synth_weights(odims, seed=seed)
    if stop_boost is not None:
        weights["lm_head.weight"][0] *= stop_boost          # the stop frame is codebook 0 == 0 (csm.py:355, 606-608)
    mcfg = omimi.MimiConfig.tiny(n_q=odims.num_codebooks, bins=odims.vocab_size)
    msd = omimi.synth_state_dict(mcfg, seed + 1)
    model = CSMModel("csm-test", state_dict={k: v.cuda() for k, v in weights.items()},
                     dims=CsmDims(**dataclasses.asdict(odims)), max_tokens=max_tokens,
                     audio_decoder_state_dict=msd, mimi_config=MimiConfig(**dataclasses.asdict(mcfg)))
    model.default_sampling_config = SamplingConfig(greedy=True)
    worker = CudaGraphWorker("csm-test", max_batch_size=max_bs, max_num_pages=pages, page_size=page, model=model,
                             max_prefill_tokens=128)
    assert isinstance(worker, DepthModelWorker) and worker.has_depth_transformer
    return worker, odims, weights, mcfg, msd


def _prompts(odims, lens, seed=7):
    g = torch.Generator().manual_seed(seed)
    N, out = odims.num_codebooks, []
    for T in lens:
        ids = torch.randint(1, odims.vocab_size, (T, N + 1), generator=g)
        ids[:, -1] = torch.randint(0, odims.text_vocab_size, (T,), generator=g)
        m = torch.zeros(T, N + 1, dtype=torch.bool)
        n_text = max(1, T // 2)
        m[:n_text, -1] = True
        m[n_text:, :N] = True
        out.append((ids, m))
    return out


def _replay(odims, w, prompt, frames, page=16, stop=0):
    """Teacher-forced oracle replay of one request under the WORKER's position rule (position T0 is skipped,
    worker/base.py:299).  Returns stats and asserts the near-tie rule."""
    ids, masks = prompt
    T0, N = ids.shape[0], odims.num_codebooks
    n_pages = (T0 + len(frames) + page - 1) // page + 1
    kv = torch.zeros(odims.num_hidden_layers, n_pages, 2, page, odims.num_key_value_heads, odims.head_dim,
                     dtype=w["lm_head.weight"].dtype)
    pages = list(range((T0 + page - 1) // page))
    pre = lm_ops.PagedWrapperCPU("prefill", page)
    pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page])
    logits, hidden = ocsm.backbone_forward(w, odims, ocsm.frame_embeds(w, odims, ids, masks),
                                           torch.arange(T0, dtype=torch.int32), pre, kv)
    logits, hidden = logits[-1], hidden[-1]
    mask = torch.ones(1, N + 1, dtype=torch.bool)
    mask[0, -1] = False
    st = dict(rows=0, flips=0)

    def check(lg, got):
        lg = lg.float()
        st["rows"] += 1
        if int(torch.argmax(lg)) != got:
            assert float(lg.max() - lg[got]) <= 2 * TOL * float(lg.abs().max()), (got, int(torch.argmax(lg)))
            st["flips"] += 1

    kv_len, pos = T0, T0 + 1
    for f, fr in enumerate(frames):
        assert fr[N] == fr[0]                                          # text column = codebook 0 (csm.py:693)
        check(logits, fr[0])
        _, dl = ocsm.depth_loop_greedy(w, odims, hidden, fr[0], forced=fr[1:N])
        for c in range(1, N):
            check(dl[c - 1], fr[c])
        if fr[0] == stop:
            assert f == len(frames) - 1
            break
        kv_len += 1
        if (kv_len + page - 1) // page > len(pages):
            pages.append(len(pages))
        dec = lm_ops.PagedWrapperCPU("decode", page)
        dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page])
        row = torch.tensor([list(fr[:N]) + [0]], dtype=torch.long)
        lg, hd = ocsm.backbone_forward(w, odims, ocsm.frame_embeds(w, odims, row, mask),
                                       torch.tensor([pos], dtype=torch.int32), dec, kv)
        logits, hidden, pos = lg[0], hd[0], pos + 1
    return st


def _serve(worker, prompts, async_mode=False, stagger=3):
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    sched = Scheduler(worker)
    reqs = [Request(request_id=f"c{i}", prompt=p) for i, p in enumerate(prompts)]
    state, pending = None, list(reqs)
    for _ in range(2000):
        if pending:
            sched.submit(pending.pop(0))              # a new request every `stagger` steps: mixed prefill + decode steps
        n = stagger if pending else None
        if async_mode:
            state = sched.run_async(n, state)
        elif n is None:
            sched.run_until_done(max_steps=3000)
        else:
            for _ in range(n):
                sched._step()
        if not pending:
            break
    torch.cuda.synchronize()
    assert not sched.has_work() and len(sched.finished) == len(reqs)
    return sched, reqs


def _check_audio(sched, reqs, mcfg, msd, N, interval=10, full_only=False):
    """full_only (async scheduling): the frame of the one extra LM step a finished request still runs reaches the host
    list after the final chunk was cut, so only the full windows are compared there."""
    n_chunks = 0
    for r in reqs:
        frames = [t[0].tolist() for t in r.lm_output_audio_tokens]
        chunks = sched.audio[r.request_id]
        want_chunks = (len(frames) + interval - 1) // interval
        if full_only:
            chunks = chunks[:len(frames) // interval]
            assert len(chunks) == len(frames) // interval
        elif r.finish_reason == "stop_id_encountered" and frames and len(frames) % interval == 0:
            # reference scheduling quirk, mirrored: the stop frame arrives after the last full window has been vocoded,
            # the request is then re-selected with its stale decode index "so _send_responses can send completion"
            # (scheduler/base.py:318-326) and run_detokenize vocodes that window once more
            assert len(chunks) == want_chunks + 1 and chunks[-1] == chunks[-2], (r.request_id, len(frames), len(chunks))
            chunks = chunks[:-1]
        else:
            assert len(chunks) == want_chunks, (r.request_id, len(frames), len(chunks))
        for ci, blob in enumerate(chunks):
            win = frames[ci * interval:(ci + 1) * interval]
            n_valid = len(win)
            win = win + [win[-1]] * (interval - n_valid)                                    # worker/base.py:629-632
            codes = torch.tensor(win)[:, :N].t()[None].clamp(0, mcfg.bins - 1)
            ref = omimi.decode(msd, mcfg, codes)[0].numpy()
            ref16 = (ref * 32767).astype(np.int16)
            if n_valid < interval:
                ref16 = ref16[:, :int(ref16.shape[1] * (n_valid - 0.5) / interval)]          # :661-668
            got = np.frombuffer(blob, dtype=np.int16).reshape(1, -1)
            assert got.shape == ref16.shape, (got.shape, ref16.shape)
            assert np.abs(got.astype(np.int32) - ref16.astype(np.int32)).max() <= 2
            n_chunks += 1
    return n_chunks


@pytest.mark.parametrize("async_mode", [False, True], ids=["sync", "async"])
def test_csm_worker_e2e_max_tokens(async_mode):
    lens = (9, 23, 5, 14)
    worker, odims, w, mcfg, msd = _build(seed=11, max_bs=4, stop_boost=0.0, max_tokens=max(lens) + 24)
    assert worker.capture_decode_graphs() >= 4
    prompts = _prompts(odims, lens)
    sched, reqs = _serve(worker, prompts, async_mode)
    N, tot = odims.num_codebooks, dict(rows=0, flips=0)
    for r, p in zip(reqs, prompts):
        frames = [t[0].tolist() for t in r.lm_output_tokens]
        assert r.finish_reason == "max_tokens_reached" and len(frames) >= 20, (r.finish_reason, len(frames))
        # async scheduling runs one more LM step after the request finished (scheduler/base.py:168-215): its frame is
        # in lm_output_tokens but conditioned on a state the host no longer tracks; the oracle replays the tracked ones
        audio = [t[0].tolist() for t in r.lm_output_audio_tokens]
        st = _replay(odims, w, p, frames[:len(audio)])
        tot["rows"] += st["rows"]
        tot["flips"] += st["flips"]
    n_chunks = _check_audio(sched, reqs, mcfg, msd, N, full_only=async_mode)
    print("csm worker e2e:", tot, "chunks", n_chunks, "steps", sched.steps, "launches", worker.gpu_launches)
    assert tot["flips"] <= max(2, tot["rows"] // 100), tot
    assert n_chunks >= 2 * len(lens) and worker.gpu_launches > 0
    assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == 4


def test_csm_worker_stop_frames_and_short_final_chunk():
    lens = (12, 6, 17)
    worker, odims, w, mcfg, msd = _build(seed=13, max_bs=3, stop_boost=1.6, max_tokens=200)
    prompts = _prompts(odims, lens)
    sched, reqs = _serve(worker, prompts)
    N, stops = odims.num_codebooks, 0
    for r, p in zip(reqs, prompts):
        frames = [t[0].tolist() for t in r.lm_output_tokens]
        if r.finish_reason == "stop_id_encountered":
            stops += 1
            assert frames[-1][0] == 0 and len(r.lm_output_audio_tokens) == len(frames) - 1
        _replay(odims, w, p, frames)
    assert stops >= 1, [r.finish_reason for r in reqs]
    _check_audio(sched, reqs, mcfg, msd, N)
    assert worker.empty_pages.qsize() == worker.max_num_pages
