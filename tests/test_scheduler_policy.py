"""Host-side scheduler policy (no GPU): window selection of Scheduler._select_detokenize_requests
(scheduler/base.py:302-333) with the default schedule and with the optional vocoder batching."""
from types import SimpleNamespace

from vox_serve_b200.requests import Request
from vox_serve_b200.scheduler import Scheduler


def _worker():
    return SimpleNamespace(max_batch_size=8, detokenize_interval=28, detokenize_overlap=21)


def _req(i, n_tokens, emitted_windows=0):
    r = Request(request_id=f"r{i}", prompt=[1, 2, 3], model_kwargs={})
    r.done_lm_prefill = True
    r.lm_output_audio_tokens = list(range(n_tokens))
    if emitted_windows:
        r.next_audio_decode_idx = [7 * (emitted_windows - 1)]
    return r


def test_default_schedule_selects_every_ready_window():
    s = Scheduler(_worker())
    s.active_requests = [_req(0, 27), _req(1, 28), _req(2, 34, 1), _req(3, 35, 1)]
    sel = s._select_detokenize_requests()
    assert [r.request_id for r in sel] == ["r1", "r3"]
    assert sel[0].next_audio_decode_idx == [0] and sel[1].next_audio_decode_idx == [7]


def test_vocoder_batching_holds_later_chunks_only():
    s = Scheduler(_worker(), vocoder_batch_steps=7)
    first, later = _req(0, 28), _req(1, 35, 1)
    done = _req(2, 30, 1)
    done.done_lm_generation = True
    s.active_requests = [first, later, done]
    held = []
    for step in range(1, 8):
        sel = [r.request_id for r in s._select_detokenize_requests()]
        if step == 1:
            assert sel == ["r0", "r2"], sel             # a first chunk and a finished request are never held
        held.append("r1" in sel)
    assert held == [False] * 6 + [True]                  # the later chunk waits for the 7th selection
    assert later.next_audio_decode_idx == [7]
