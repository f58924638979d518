"""Qwen3-TTS served end to end through the worker API (BASELINE.json configs[2]'s path):
scheduler loop -> DepthModelWorker -> Qwen3TTSModel.frame_device (talker + code predictor, one device-side frame per step,
input_features carried per slot) -> the streaming 12 Hz codec decoder with ONE Qwen3TTSDecoderCache over the batch slots.

The CPU oracle replays every request on its own, teacher-forced with the GPU's frames: every codebook decision must be the
oracle's argmax unless its own top-2 margin is a bf16 near-tie; every audio chunk must equal the oracle codec's forward_chunk
on the same frames, chunk after chunk with the request's own cache (int16, +-2 LSB)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import lm_ops, qwen3_codec as ocodec, qwen3_tts as oq

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _build(seed, max_bs, max_tokens, pages=64, page=16):
    from vox_serve_b200.depth_engine import Qwen3TTSDims
    from vox_serve_b200.model.qwen3_tts import Qwen3TTSModel
    from vox_serve_b200.sampling import SamplingConfig
    from vox_serve_b200.tokenizer.qwen3_codec import Qwen3CodecConfig
    from vox_serve_b200.worker import CudaGraphWorker, DepthModelWorker

    od = oq.Qwen3TTSDims.tiny()
    weights = oq.synth_weights(od, seed=seed)
    ccfg = ocodec.Qwen3CodecConfig.tiny(num_quantizers=od.num_code_groups, codebook_size=od.cp_vocab_size)
    csd = ocodec.synth_state_dict(ccfg, seed + 1)
    model = Qwen3TTSModel("qwen3-test", state_dict={k: v.cuda() for k, v in weights.items()},
                          dims=Qwen3TTSDims(**dataclasses.asdict(od)), max_tokens=max_tokens, audio_decoder_state_dict=csd,
                          codec_config=Qwen3CodecConfig(**dataclasses.asdict(ccfg)), stop_token_id=-1)
    model.default_sampling_config = SamplingConfig(greedy=True)
    worker = CudaGraphWorker("qwen3-test", max_batch_size=max_bs, max_num_pages=pages, page_size=page, model=model,
                             max_prefill_tokens=128)
    assert isinstance(worker, DepthModelWorker) and worker.feats is not None and worker.voc_cache is not None
    return worker, od, weights, ccfg, csd


def _prompts(od, lens, seed=7):
    g = torch.Generator().manual_seed(seed)
    N, out = od.num_code_groups, []
    for T in lens:
        ids = torch.zeros(T, N + 1, dtype=torch.int64)
        ids[:, -1] = torch.randint(0, od.text_vocab_size, (T,), generator=g)
        ids[:, 0] = torch.randint(0, od.vocab_size, (T,), generator=g)
        m = torch.zeros(T, N + 1, dtype=torch.bool)
        m[T // 2:, -1] = True                               # the second half of the prompt carries codec embeddings
        feats = (torch.randn(T, od.hidden_size, generator=g) * 0.5).to(torch.bfloat16)
        out.append((ids, m, feats))
    return out


def _replay(od, w, prompt, frames, page=16):
    ids, masks, feats = prompt
    T0, N = ids.shape[0], od.num_code_groups
    n_pages = (T0 + len(frames) + page - 1) // page + 1
    kv = torch.zeros(od.num_hidden_layers, n_pages, 2, page, od.num_key_value_heads, od.head_dim, dtype=torch.bfloat16)
    pages = list(range((T0 + page - 1) // page))
    pre = lm_ops.PagedWrapperCPU("prefill", page)
    pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page])
    logits, hidden = oq.talker_forward(w, od, oq.talker_embeds(w, ids[:, -1], ids[:, 0], masks[:, -1], feats),
                                       torch.arange(T0, dtype=torch.int32), pre, kv)
    logits, hidden = logits[-1], hidden[-1]
    st = dict(rows=0, flips=0)

    def check(lg, got):
        lg = lg.float()
        st["rows"] += 1
        if int(torch.argmax(lg)) != got:
            assert float(lg.max() - lg[got]) <= 2 * TOL * float(lg.abs().max()), (got, int(torch.argmax(lg)))
            st["flips"] += 1

    kv_len, pos = T0, T0 + 1                                 # position T0 is skipped (worker/base.py:299)
    for fr in frames:
        assert fr[N] == od.tts_pad_token_id                  # text column of a generated frame (qwen3_tts.py:1916)
        check(logits, fr[0])
        _, cl, feat = oq.predictor_loop_greedy(w, od, hidden, fr[0], forced=fr[1:N])
        for c in range(1, N):
            check(cl[c - 1], fr[c])
        kv_len += 1
        if (kv_len + page - 1) // page > len(pages):
            pages.append(len(pages))
        dec = lm_ops.PagedWrapperCPU("decode", page)
        dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page])
        e = oq.talker_embeds(w, torch.tensor([od.tts_pad_token_id]), torch.tensor([fr[0]]), torch.tensor([True]), feat)
        lg, hd = oq.talker_forward(w, od, e, torch.tensor([pos], dtype=torch.int32), dec, kv)
        logits, hidden, pos = lg[0], hd[0], pos + 1
    return st


@pytest.mark.parametrize("async_mode", [False, True], ids=["sync", "async"])
def test_qwen3_tts_worker_e2e(async_mode):
    _serve_and_check((9, 21, 5), 3, async_mode)


def test_qwen3_tts_worker_recycles_slots_and_codec_state():
    """More requests than batch slots: a stream that starts on a recycled slot must see zeroed codec state (attention
    window, position offset, every conv cache) and its own input_features -- its audio is compared with the oracle codec
    started from a fresh cache."""
    _serve_and_check((7, 12, 5, 9, 6), 2, False)


def _serve_and_check(lens, max_bs, async_mode):
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    worker, od, w, ccfg, csd = _build(seed=11, max_bs=max_bs, max_tokens=max(lens) + 26)
    prompts = _prompts(od, lens)
    sched = Scheduler(worker)
    reqs = [Request(request_id=f"q{i}", prompt=p) for i, p in enumerate(prompts)]
    state, pending = None, list(reqs)
    while pending:
        sched.submit(pending.pop(0))                          # a new request every 3 steps: mixed prefill + decode steps
        if async_mode:
            state = sched.run_async(3 if pending else None, state)
        elif pending:
            for _ in range(3):
                sched._step()
        else:
            sched.run_until_done(max_steps=3000)
    torch.cuda.synchronize()
    assert not sched.has_work() and len(sched.finished) == len(reqs)
    N, interval, tot = od.num_code_groups, 10, dict(rows=0, flips=0)
    n_chunks = 0
    for r, p in zip(reqs, prompts):
        audio = [t[0].tolist() for t in r.lm_output_audio_tokens]
        assert r.finish_reason == "max_tokens_reached" and len(audio) >= 20
        st = _replay(od, w, p, audio)
        tot["rows"] += st["rows"]
        tot["flips"] += st["flips"]
        # audio: chunk after chunk through the oracle codec with this request's own cache
        chunks = sched.audio[r.request_id]
        n_full = len(audio) // interval
        if not async_mode:
            assert len(chunks) == (len(audio) + interval - 1) // interval
        cache = ocodec.init_cache(ccfg, 1)
        for ci in range(len(chunks) if not async_mode else n_full):
            win = audio[ci * interval:(ci + 1) * interval]
            n_valid = len(win)
            win = win + [win[-1]] * (interval - n_valid)
            codes = torch.tensor(win)[:, :N].t()[None].clamp(0, ccfg.codebook_size - 1)
            ref, cache = ocodec.forward_chunk(csd, ccfg, codes, cache)
            ref16 = (ref[0].numpy() * 32767).astype(np.int16)
            if n_valid < interval:
                ref16 = ref16[:, :int(ref16.shape[1] * (n_valid - 0.5) / interval)]
            got = np.frombuffer(chunks[ci], dtype=np.int16).reshape(1, -1)
            assert got.shape == ref16.shape, (got.shape, ref16.shape)
            assert np.abs(got.astype(np.int32) - ref16.astype(np.int32)).max() <= 2, (r.request_id, ci)
            n_chunks += 1
    print("qwen3-tts worker e2e:", tot, "chunks", n_chunks, "steps", sched.steps, "launches", worker.gpu_launches)
    assert tot["flips"] <= max(2, tot["rows"] // 100), tot
    assert n_chunks >= 2 * len(lens)
    assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == max_bs
