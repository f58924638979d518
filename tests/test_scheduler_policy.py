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


def test_batched_prefill_selection_packs_prompts_by_length():
    """Scheduler(one_prefill_per_step=False) (scheduler/base.py:283-286 made selectable): prompts whose length is known
    (after their first preprocess) are packed up to the worker's prefill row budget; decode requests fill the rest."""
    w = SimpleNamespace(max_batch_size=8, detokenize_interval=28, detokenize_overlap=21, prefill_graph_batch_size=8,
                        cuda_graph_seq_len_buckets=[300])
    s = Scheduler(w, one_prefill_per_step=False)
    waiting = []
    for i, n in enumerate((133, 133, 133, 20)):
        r = Request(request_id=f"p{i}", prompt=[0] * n, model_kwargs={})
        r.input_length = n
        waiting.append(r)
    running = _req(9, 30, 1)
    s.active_requests = waiting + [running]
    sel = [r.request_id for r in s._select_lm_requests()]
    assert sel == ["p0", "p1", "p3", "r9"], sel           # 133 + 133 + 20 <= 300, the third 133 does not fit
    s1 = Scheduler(w)                                      # the reference's default: one prefill per step
    s1.active_requests = waiting + [running]
    assert [r.request_id for r in s1._select_lm_requests()] == ["p0", "r9"]


def test_detach_and_adopt_move_a_request_between_loops():
    """Scheduler.detach / adopt: the scheduler half of a KV hand-off (vox_serve_b200/kv_handoff.py)."""
    import pytest

    a, b = Scheduler(_worker()), Scheduler(_worker())
    r0, r1 = _req(0, 30, 1), _req(1, 10)
    a.submit(r0)
    a.submit(r1)
    a._prepare_requests()
    a.audio["r0"].append(b"x")
    moved = a.detach("r0")
    assert moved is r0 and [r.request_id for r in a.active_requests] == ["r1"]
    assert a.audio["r0"] == [b"x"]                         # what was delivered before the hand-off stays accounted for
    with pytest.raises(KeyError):
        a.detach("r0")
    b.adopt(moved)
    assert b.active_requests == [r0] and b.audio["r0"] == [] and "r0" in b.submit_time
    assert b.has_work()
    sel = b._select_detokenize_requests()                  # the carried vocoder progress decides the next window
    assert sel == [] and r0.next_audio_decode_idx == [0]   # 30 tokens: the window at 7 needs 35
