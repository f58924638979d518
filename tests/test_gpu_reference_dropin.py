"""The drop-in boundary exercised by the REFERENCE'S OWN CODE.

The unmodified reference package (``pip install --target baseline/_ref /root/reference``; git-ignored, shipped to
the GPU box with the snapshot) is imported with INTEGRATION.md's ``sys.modules`` redirection applied:

* ``test_reference_orpheus_adapter_runs_on_b200_operator_shims``: the reference's ``OrpheusForCausalLM`` /
  ``OrpheusModel.forward`` / ``OrpheusModel.sampling`` (model/orpheus.py:41-221, 398-477; nn.Linear on cuBLAS) call
  OUR ``FlashInferPrefillWrapper/DecodeWrapper.plan(host int32 tensors) -> set_kv_cache(kv_cache[i], k, v) ->
  run(q, kv_cache[i])``, ``rms_norm``, ``apply_rope_pos_ids(**llama-3.1 kwargs)`` and ``Sampler.apply_repetition_penalty /
  run_sampling / update_repetition_penalty_cache`` exactly as the reference worker drives them
  (worker/base.py:396-520); results are held to the CPU oracle and to ``LlamaEngine`` on the same weights.
* ``test_reference_scheduler_drives_b200_worker_over_zmq``: the reference's ``Scheduler`` (scheduler/base.py:14-478)
  constructs ``vox_serve_b200.worker.CudaGraphWorker`` by name through its own ``worker_kwargs`` and serves three
  requests pushed through its real ZMQ request socket; audio + completion messages come back through its result socket.
"""
import json
import os
import time

import numpy as np
import pytest
import torch

from oracle import orpheus as oorph, sampler as osampler, worker as oworker
from tests import ref_dropin

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_dropin.available(), reason="baseline/_ref (pip-installed reference) is absent")]
BF = torch.bfloat16


def test_reference_orpheus_adapter_runs_on_b200_operator_shims():
    ref_dropin.install_redirects()
    import vox_serve.model.orpheus as ro
    from vox_serve.flashinfer_utils import FlashInferDecodeWrapper, FlashInferPrefillWrapper
    from vox_serve.requests import Request
    from vox_serve.sampling import SamplingConfig

    import vox_serve_b200.flashinfer_utils as ours
    from vox_serve_b200 import ops
    from vox_serve_b200.engine import LlamaDims, LlamaEngine, LlamaWeights

    assert ro.rms_norm is ours.rms_norm and ro.apply_rope_pos_ids is ours.apply_rope_pos_ids
    assert FlashInferDecodeWrapper is ours.FlashInferDecodeWrapper and "baseline/_ref" in ro.__file__

    dims = oorph.OrpheusDims.tiny(hidden_size=1536, num_hidden_layers=3, num_attention_heads=12, num_key_value_heads=4,
                                  intermediate_size=2048, vocab_size=10 + 7 * 4096)
    dims.max_tokens = 400
    page, pages, max_bs = 128, 16, 4
    weights = oorph.synth_weights(dims, seed=4)
    lm = ro.OrpheusForCausalLM(ref_dropin.llama_config(dims))
    lm.load_state_dict(weights, strict=True)
    lm = lm.to(BF).cuda().eval()
    cfg = SamplingConfig(top_k=None, top_p=0.8, min_p=None, temperature=0.6, repetition_penalty=1.3, repetition_window=-1,
                         cfg_scale=None, greedy=True)
    m = object.__new__(ro.OrpheusModel)          # skip the hub download of OrpheusModel.__init__ (orpheus.py:238-250)
    m.model_name, m.device, m.dtype, m.model = "synthetic", "cuda:0", BF, lm
    m.stop_token_id, m.default_sampling_config = dims.stop_token_id, cfg

    # the wrappers exactly as the reference worker constructs them (worker/base.py:149-166)
    buf = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    kw = dict(attn_buffer=buf, n_qo_head=dims.num_attention_heads, n_kv_head=dims.num_key_value_heads,
              n_state=dims.num_attention_heads * dims.head_dim, page_size=page, use_cuda_graph=False)
    prefill_wrapper, decode_wrapper = FlashInferPrefillWrapper(**kw), FlashInferDecodeWrapper(**kw)
    kv_cache = torch.zeros(dims.num_hidden_layers, pages, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF,
                           device="cuda")
    # our engine on the same weights and its own cache, as the second reference point
    ld = LlamaDims(dims.hidden_size, dims.num_hidden_layers, dims.num_attention_heads, dims.num_key_value_heads, dims.head_dim,
                   dims.intermediate_size, dims.vocab_size, dims.rms_norm_eps, dims.rope_theta, dims.rope_factor,
                   dims.low_freq_factor, dims.high_freq_factor, dims.old_context_len)
    eng = LlamaEngine(LlamaWeights.from_state_dict(weights, ld), torch.zeros_like(kv_cache), page, max_rows=512)

    ocfg = osampler.SamplingConfig(top_p=0.8, temperature=0.6, repetition_penalty=1.3, repetition_window=-1, greedy=True,
                                   max_tokens=dims.max_tokens)
    ow = oworker.OracleWorker(weights, dims, ocfg, page_size=page, max_num_pages=pages, max_batch_size=max_bs)
    g = torch.Generator().manual_seed(21)
    prompt_lens = [133, 120, 7]
    oreqs = [oworker.Req(f"r{i}", torch.randint(10, dims.vocab_size, (n,), generator=g)) for i, n in enumerate(prompt_lens)]
    rreqs = {r.request_id: Request(request_id=r.request_id, prompt=None) for r in oreqs}
    active = []
    st = dict(rows=0, id_mismatch=0, err_oracle=0.0, err_engine=0.0)
    for step in range(20):
        if step < len(oreqs):
            active.append(oreqs[step])
        lmr = ow.select_lm(active)
        inp = ow.prepare_lm_inputs(lmr)
        # ---- the reference worker's step (worker/base.py:396-520), reference adapter, our operators ----
        i32 = lambda x: torch.tensor(x, dtype=torch.int32)       # noqa: E731  (host tensors, as the worker builds them)
        if inp["is_prefill"]:
            wrapper = prefill_wrapper
            wrapper.plan(i32(inp["qo_indptr"]), i32(inp["paged_kv_indptr"]), i32(inp["paged_kv_indices"]),
                         i32(inp["paged_kv_last_page_len"]), torch.bfloat16)
        else:
            wrapper = decode_wrapper
            wrapper.plan(i32(inp["paged_kv_indptr"]), i32(inp["paged_kv_indices"]), i32(inp["paged_kv_last_page_len"]),
                         torch.bfloat16)
        torch.cuda.synchronize()
        input_ids, position_ids = inp["input_ids"].to(torch.int32).cuda(), inp["position_ids"].cuda()
        with torch.no_grad():
            logits = m.forward(input_ids=input_ids, position_ids=position_ids, attn_wrapper=wrapper, kv_cache=kv_cache,
                               input_features=None, input_masks=None)
        if inp["is_prefill"]:
            logits = logits[wrapper.qo_indptr[1:].long() - 1]        # cuda_graph_worker.py:900-902
        assert logits.shape == (len(lmr), 1, dims.vocab_size) and logits.dtype == BF
        rep = inp["repetition_cache"].cuda()
        for r in lmr:                                  # what the reference's prepare_lm_inputs maintains (base.py:299, 325)
            rreqs[r.request_id].next_position_id = r.next_position_id
        ids, task = m.sampling(logits=logits, requests=[rreqs[r.request_id] for r in lmr], repetition_cache=rep)
        import asyncio
        asyncio.run(task)
        ids = ids.cpu()
        # ---- our engine on the same step ----
        R = input_ids.shape[0]
        d = lambda x: torch.tensor(x, dtype=torch.int32, device="cuda")   # noqa: E731
        ops.plan_rows(eng.plan, d(inp["qo_indptr"]) if inp["is_prefill"] else None, d(inp["paged_kv_indptr"]),
                      d(inp["paged_kv_indices"]), d(inp["paged_kv_last_page_len"]), len(lmr), R, page, eng.chunk)
        last_rows = d([x - 1 for x in inp["qo_indptr"][1:]]) if inp["is_prefill"] else None
        e_logits = eng.forward(input_ids[:, 0].contiguous(), position_ids, R, last_rows=last_rows).float().cpu()
        # ---- oracle, teacher-forced with the reference adapter's ids ----
        rep_before = inp["repetition_cache"].clone()
        ow.run_lm(lmr, inp, forced_ids=ids.view(-1, 1).to(torch.int64))
        ref_logits = ow.last_logits[:, 0].float()
        got = logits[:, 0].float().cpu()
        scale = float(ref_logits.abs().max())
        st["err_oracle"] = max(st["err_oracle"], float((got - ref_logits).abs().max()) / scale)
        st["err_engine"] = max(st["err_engine"], float((got - e_logits).abs().max()) / scale)
        # the repetition cache the reference adapter updated through OUR Sampler == the oracle's update (batch-union rule)
        assert torch.equal(rep.cpu(), inp["repetition_cache"]), step
        pen = osampler.apply_repetition_penalty(ow.last_logits, rep_before, 1.3)[:, 0].float()
        top2 = torch.topk(pen, 2, dim=-1).values
        for r in range(len(lmr)):
            st["rows"] += 1
            if int(ids[r, 0]) != int(ow.last_own_ids[r, 0]):
                st["id_mismatch"] += 1
                assert float(top2[r, 0] - top2[r, 1]) <= 4 * float(top2[r, 0].abs()) * 2.0 ** -8, (step, r)
            # request state written by the reference's update_req_states through our ids
            assert int(rreqs[lmr[r].request_id].lm_output_tokens[-1][0, 0]) == int(ids[r, 0])
    print("reference adapter on B200 shims:", st)
    assert st["err_oracle"] < 2e-2 and st["err_engine"] < 2e-2, st
    assert st["id_mismatch"] <= 2, st


def test_reference_scheduler_drives_b200_worker_over_zmq(tmp_path):
    import zmq

    ref_dropin.install_redirects(worker=True)
    import vox_serve.scheduler.base as sb

    import vox_serve_b200.worker as ours

    assert sb.CudaGraphWorker is ours.CudaGraphWorker and "baseline/_ref" in sb.__file__
    req_path, res_path = str(tmp_path / "req.ipc"), str(tmp_path / "res.ipc")
    ctx = zmq.Context()
    results = ctx.socket(zmq.PULL)
    results.bind(f"ipc://{res_path}")
    sched = sb.Scheduler(model_name_or_path="orpheus-synthetic-tiny:3", max_batch_size=4, max_num_pages=64, page_size=16,
                         request_socket_path=req_path, result_socket_path=res_path, greedy=True, max_tokens=90)
    assert isinstance(sched.model_worker, ours.CudaGraphWorker)
    push = ctx.socket(zmq.PUSH)
    push.connect(f"ipc://{req_path}")
    g = torch.Generator().manual_seed(21)
    prompts = {f"req{i}": torch.randint(10, 128000, (n,), generator=g).tolist() for i, n in enumerate((9, 22, 40))}
    for rid, p in prompts.items():
        msg = {"request_id": rid, "prompt": p, "is_streaming": True, "model_kwargs": {"voice": None}}
        push.send(json.dumps(msg).encode("utf-8") + b"|")
    time.sleep(0.2)
    audio, done = {rid: [] for rid in prompts}, {}
    for _ in range(600):
        sched._step()                                   # the reference's own loop body (run_forever, :223-232)
        torch.cuda.synchronize()
        while True:
            try:
                payload = results.recv(flags=zmq.NOBLOCK)
            except zmq.Again:
                break
            rid, kind, body = payload.split(b"|", 2)
            if kind == b"AUDIO":
                audio[rid.decode()].append(body)
            else:
                done[rid.decode()] = json.loads(body.decode())
        if len(done) == len(prompts):
            break
    assert set(done) == set(prompts) and all(v["status"] == "completed" for v in done.values()), done
    assert all(v["reason"] == "max_tokens_reached" for v in done.values()), done
    w = sched.model_worker
    assert w.empty_pages.qsize() == w.max_num_pages and len(w.free_slots) == w.max_batch_size
    for rid, p in prompts.items():
        n_prompt = len(p) + 5
        n_tok = 90 - n_prompt                 # next_position_id > max_tokens stops the request (orpheus.py:468-471)
        assert len(audio[rid]) >= 1
        samples = sum(len(c) for c in audio[rid]) // 2
        full = max(0, (n_tok - 28) // 7 + 1)
        assert samples >= 2048 * full and all(len(c) % 2 == 0 for c in audio[rid]), (rid, samples, full)
        pcm = np.frombuffer(b"".join(audio[rid]), dtype=np.int16)
        assert np.abs(pcm).max() > 0
    push.close(0), results.close(0), sched.request_socket.close(0), sched.result_socket.close(0)
