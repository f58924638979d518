"""Worker-level end-to-end parity on the GPU: prefill + decode + SNAC + PCM through the reference-facing worker
API (prepare_lm_inputs / run_lm_prefill / run_lm_decode / run_detokenize / free_kv_cache) against the CPU oracle,
teacher-forced (see tests/e2e_harness.py)."""
import pytest

from oracle import orpheus as oorph

pytestmark = pytest.mark.gpu


def _check(st):
    print(st)
    # every disagreement must be a bf16 near-tie, and there must be few of them
    assert st["id_mismatch"] == st["low_margin"], st
    assert st["id_mismatch"] <= max(2, st["rows"] // 40), st
    # identical tokens + identical noise -> PCM within a few int16 steps (waveform tolerance 1e-3 relative
    # = 32 steps at full scale; fp32 reassociation in the conv stack stays far below that)
    assert st["chunks"] > 0 and st["pcm_max_lsb"] <= 4, st
    assert st["gpu_launches"] > 0


def test_e2e_tiny_ragged_batch():
    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    _check(__import__("tests.e2e_harness", fromlist=["x"]).run_e2e_parity(
        prompt_lens=(9, 16, 30, 33), n_tokens=45, seed=3, dims=dims, page_size=16, max_pages=128))


def test_e2e_page128_gqa3():
    # Orpheus attention geometry (head_dim 128, 3 q heads per kv head, page 128) with a prompt crossing a page
    dims = oorph.OrpheusDims.tiny(hidden_size=1536, num_hidden_layers=3, num_attention_heads=12,
                                  num_key_value_heads=4, intermediate_size=2048, vocab_size=156940,
                                  stop_token_id=128258, audio_id_base=128266)
    _check(__import__("tests.e2e_harness", fromlist=["x"]).run_e2e_parity(
        prompt_lens=(133, 120, 12), n_tokens=36, seed=4, dims=dims, page_size=128, max_pages=16))


def test_async_scheduling_matches_sync_schedule_lengths():
    """Scheduler._step_async ordering (host state one step behind the device): every request still produces the
    same number of tokens and audio chunks as under the synchronous loop, and PCM sizes agree."""
    import torch

    from oracle import snac as osnac
    from tests.e2e_harness import build_models
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    prompt_lens, n_tokens = (9, 16, 30, 33), 45
    dims.max_tokens = max(prompt_lens) + n_tokens
    worker, _ = build_models(dims, osnac.SnacConfig.tiny(), 3, len(prompt_lens), 16, 128)
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist() for n in prompt_lens]
    out = {}
    for mode in ("sync", "async"):
        sched = Scheduler(worker)
        reqs = [Request(request_id=f"{mode}{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(prompts)]
        for r in reqs:
            sched.submit(r)
        if mode == "sync":
            sched.run_until_done(max_steps=4000)
        else:
            sched.run_async()
        torch.cuda.synchronize()
        assert all(r.done_all for r in reqs), mode
        out[mode] = [(len(r.lm_output_audio_tokens), [len(c) for c in sched.audio[r.request_id]], r.finish_reason)
                     for r in reqs]
        assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == worker.max_batch_size
    # the async loop may run one extra LM step per request before it notices max_tokens (reference behaviour)
    for (ns, cs, fs), (na, ca, fa) in zip(out["sync"], out["async"]):
        assert fs == fa and 0 <= na - ns <= 1, (ns, na)
        assert abs(len(ca) - len(cs)) <= 1 and ca[:len(cs) - 1] == cs[:len(cs) - 1]


def test_resident_loop_matches_per_step_api():
    """ModelWorker.run_lm_decode_resident (device-resident CUDA-graph loop, vocoder graph overlapped on the side
    stream) against the per-step worker API: identical greedy tokens for every request, and the PCM it leaves in
    HBM equals the eager vocoder pass over each request's newest 28-token window (same injected noise)."""
    import torch

    from oracle import snac as osnac
    from tests.e2e_harness import build_models
    from vox_serve_b200 import ops
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    prompt_lens, warm, span = (9, 16, 30, 33), 12, 21
    dims.max_tokens = max(prompt_lens) + 80
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist() for n in prompt_lens]
    tokens = {}
    for mode in ("resident", "per_step"):
        worker, _ = build_models(dims, osnac.SnacConfig.tiny(), 3, len(prompt_lens), 16, 128)
        cache = {}

        def fixed_noise(shapes, cache=cache):
            for s in shapes:
                if tuple(s) not in cache:
                    cache[tuple(s)] = torch.randn(s, generator=torch.Generator().manual_seed(sum(s))).cuda()
            return [cache[tuple(s)] for s in shapes]

        worker.model.audio_decoder.noise_source = fixed_noise
        sched = Scheduler(worker)
        reqs = [Request(request_id=f"{mode}{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(prompts)]
        for r in reqs:
            sched.submit(r)
        for _ in range(warm):
            sched._step()
        assert all(r.done_lm_prefill and 7 <= len(r.lm_output_audio_tokens) < 28 for r in reqs)
        if mode == "per_step":
            for _ in range(span):
                sched._step()
        else:
            before = [len(r.lm_output_audio_tokens) for r in reqs]
            # (the noise tensors must exist before the vocoder graph is captured: no host copies inside a capture)
            fixed_noise(worker.model.audio_decoder.noise_shapes(len(reqs), 16))
            worker.run_lm_decode_resident(reqs, span, detokenize=True)
            torch.cuda.synchronize()
            assert [len(r.lm_output_audio_tokens) for r in reqs] == [b + span for b in before]
            B, W = len(reqs), worker.detokenize_interval
            last = torch.tensor([[int(t) for t in r.lm_output_audio_tokens[-W:]] for r in reqs], dtype=torch.int64).cuda()
            want = ops.pcm16(worker.model.postprocess(last.view(B, W, 1)))
            torch.cuda.synchronize()
            assert torch.equal(worker.res_pcm[:B].cpu(), want.cpu())
            assert worker.res_pcm[:B].abs().max().item() > 0
        tokens[mode] = [[int(t) for t in r.lm_output_audio_tokens] for r in reqs]
    assert tokens["resident"] == tokens["per_step"]


def test_e2e_stop_token_sync():
    """Stop id NOT masked (its lm_head row is boosted so it wins the argmax after 10-50 tokens): the
    'stop_id_encountered' path, the padded last window with its (n_valid - 0.5) / interval trim, max_tokens for the
    request that never stops, and slot / page release -- against the oracle worker, teacher-forced."""
    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    st = __import__("tests.e2e_harness", fromlist=["x"]).run_e2e_parity(
        prompt_lens=(9, 16, 30, 33), n_tokens=75, seed=3, dims=dims, page_size=16, max_pages=128, stop_boost=2.5)
    _check(st)
    assert st["finish_reasons"].count("stop_id_encountered") >= 2, st
    assert any(0 < n < 28 for n in st["n_audio_tokens"]), st           # a request shorter than one window
    assert any(n > 28 and n % 7 != 0 for n in st["n_audio_tokens"]), st  # a stop in the middle of a hop


@pytest.mark.parametrize("mode", ["sync", "async"])
def test_stop_token_ring_matches_host_list(mode):
    """The device-side token ring that run_detokenize reads must stay index-for-index equal to
    req.lm_output_audio_tokens when a stop id is sampled -- also when the scheduler runs one step ahead and the
    stopped request gets one more LM step (scheduler/base.py:168-215).  Every window the worker vocodes is recorded
    from the HOST token list at the moment run_detokenize is called (what the reference gathers,
    cuda_graph_worker.py:1176-1190); the delivered chunk must be the vocoder's output for exactly those tokens, byte for
    byte (zero NoiseBlock noise)."""
    import numpy as np
    import torch

    from oracle import snac as osnac
    from tests.e2e_harness import build_models
    from vox_serve_b200 import ops
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    prompt_lens = (9, 16, 30, 33)
    dims.max_tokens = max(prompt_lens) + 75
    worker, _ = build_models(dims, osnac.SnacConfig.tiny(), 3, len(prompt_lens), 16, 128, stop_boost=2.5)
    worker.model.audio_decoder.noise_source = lambda shapes: [torch.zeros(s, device="cuda") for s in shapes]
    W = worker.detokenize_interval
    snapshots = {}
    inner = worker.run_detokenize

    def recording_run_detokenize(requests):
        for r in requests:
            for d in r.audio_decode_idx:
                win = [int(t[0, 0]) for t in r.lm_output_audio_tokens[d:d + W]]
                if win:
                    snapshots.setdefault(r.request_id, []).append(win)
        return inner(requests)

    worker.run_detokenize = recording_run_detokenize
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist() for n in prompt_lens]
    sched = Scheduler(worker)
    reqs = [Request(request_id=f"s{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(prompts)]
    for r in reqs:
        sched.submit(r)
    if mode == "sync":
        sched.run_until_done(max_steps=4000)
    else:
        sched.run_async()
    torch.cuda.synchronize()
    assert all(r.done_all for r in reqs)
    assert [r.finish_reason for r in reqs].count("stop_id_encountered") >= 2
    assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == worker.max_batch_size
    n_checked = n_padded = 0
    for r in reqs:
        assert dims.stop_token_id not in [int(t[0, 0]) for t in r.lm_output_audio_tokens]
        got, wins = sched.audio[r.request_id], snapshots.get(r.request_id, [])
        assert len(got) == len(wins), (r.request_id, len(got), len(wins))
        for a, win in zip(got, wins):
            n_valid = len(win)
            ids = torch.tensor(win + [win[-1]] * (W - n_valid), dtype=torch.int64, device="cuda")
            a16 = ops.pcm16(worker.model.postprocess(ids.view(1, W, 1)))[0].cpu().numpy()
            if n_valid < W:
                a16 = a16[:, :int(a16.shape[1] * (n_valid - 0.5) / W)]
                n_padded += 1
            b = a16.tobytes()
            assert a == b, (r.request_id, len(a), len(b),
                            int(np.abs(np.frombuffer(a, np.int16).astype(int) - np.frombuffer(b, np.int16).astype(int)).max())
                            if len(a) == len(b) else -1)
            n_checked += 1
    assert n_checked >= 6 and n_padded >= 2


def test_capacity_errors_finish_the_request_not_the_loop():
    """A prompt longer than the worker accepts and a request that runs out of KV pages end with an ``error: ...``
    finish reason and give back their slot / pages; the other streams keep running and complete normally.  (The
    reference raises queue.Empty / index errors out of prepare_lm_inputs with half-updated state, worker/base.py:237-249.)"""
    import torch

    from oracle import snac as osnac
    from tests.e2e_harness import build_models
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    dims.max_tokens = 120
    # 9 pages of 16 tokens: r0 (20-token prompt -> 2 pages, grows to 8) and r2 (12 -> 1 page) cannot both run to 120
    worker, _ = build_models(dims, osnac.SnacConfig.tiny(), 3, 4, 16, 9)
    worker.max_prefill_tokens = 64
    g = torch.Generator().manual_seed(5)
    mk = lambda n: torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist()   # noqa: E731
    reqs = [Request(request_id="ok", prompt=mk(20), model_kwargs={"voice": None}),
            Request(request_id="too_long", prompt=mk(100), model_kwargs={"voice": None}),
            Request(request_id="starved", prompt=mk(12), model_kwargs={"voice": None})]
    sched = Scheduler(worker)
    for r in reqs:
        sched.submit(r)
    n = sched.run_until_done(max_steps=2000)
    torch.cuda.synchronize()
    by = {r.request_id: r for r in reqs}
    assert n < 2000 and all(r.done_all for r in reqs)
    assert by["too_long"].finish_reason.startswith("error: prompt of 100 tokens") and not sched.audio["too_long"]
    reasons = {by["ok"].finish_reason, by["starved"].finish_reason}
    assert any(x.startswith("error: out of KV pages") for x in reasons) and "max_tokens_reached" in reasons, reasons
    winner = "ok" if by["ok"].finish_reason == "max_tokens_reached" else "starved"
    assert len(sched.audio[winner]) >= 10
    assert {r.request_id for r in sched.finished} == {"ok", "too_long", "starved"}
    assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == worker.max_batch_size


def test_batched_prefill_defers_prompts_that_do_not_fit_the_step():
    """Scheduler(one_prefill_per_step=False) selects every waiting prompt (their lengths are unknown before preprocess):
    the worker takes as many as fit max_prefill_tokens rows and leaves the rest un-prefilled for the next step -- nobody
    fails, and every request generates exactly the tokens of the one-prefill-per-step schedule."""
    import torch

    from oracle import snac as osnac
    from tests.e2e_harness import build_models
    from vox_serve_b200.requests import Request
    from vox_serve_b200.scheduler import Scheduler

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    prompt_lens, n_tokens = (100, 90, 110, 70, 12), 30
    dims.max_tokens = max(prompt_lens) + n_tokens
    worker, _ = build_models(dims, osnac.SnacConfig.tiny(), 3, len(prompt_lens), 16, 256, planted=2.0)   # 256-row prefill steps
    g = torch.Generator().manual_seed(21)
    prompts = [torch.randint(10, dims.vocab_size, (n - 5,), generator=g).tolist() for n in prompt_lens]
    out = {}
    for mode, one in (("single", True), ("batched", False)):
        sched = Scheduler(worker, one_prefill_per_step=one)
        sched.trace = []
        reqs = [Request(request_id=f"{mode}{i}", prompt=p, model_kwargs={"voice": None}) for i, p in enumerate(prompts)]
        for r in reqs:
            sched.submit(r)
        sched.run_until_done(max_steps=4000)
        torch.cuda.synchronize()
        assert all(r.finish_reason == "max_tokens_reached" for r in reqs), [r.finish_reason for r in reqs]
        out[mode] = ([[int(t[0, 0]) for t in r.lm_output_tokens] for r in reqs], sched.trace)
        assert worker.empty_pages.qsize() == worker.max_num_pages and len(worker.free_slots) == worker.max_batch_size
    assert out["single"][0] == out["batched"][0]
    first = [rid for rid, _ in out["batched"][1][0]]
    # 100 + 90 rows fit the 256-row step (+ one row reserved per remaining request), 110 more do not, the 70-row prompt after
    # it still does; the 12-row one no longer
    assert first == ["batched0", "batched1", "batched3"], first
    assert len(out["batched"][1]) < len(out["single"][1])    # fewer scheduler steps overall
