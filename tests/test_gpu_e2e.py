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
