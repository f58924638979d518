"""CSM backbone + depth-transformer frames on the GPU (vox_serve/model/csm.py:158-312,
637-769; worker/cuda_graph_worker.py:1058-1160) against

* the golden file produced by the reference's own CsmBackboneModel / CsmDepthDecoderForCausalLM / CsmCodebooksHead on
  CPU (tests/golden/csm_tiny_frames.npz): ids and all logits of 5 frames x 8 codebooks, single request;
* the CPU oracle (oracle/csm.py, pinned bit-exactly to that golden) on a ragged BATCH at a mid-sized configuration
  with the production head geometry (backbone head_dim 64, depth head_dim 128, GQA 4), as one CUDA graph per frame.
Greedy; ids must match except provable near-ties, after which the GPU frame is overwritten with the oracle's so that
one flip does not fork the comparison (teacher forcing)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import csm as ocsm

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _engine(odims, weights, page, pages, max_batch, max_rows):
    from vox_serve_b200.depth_engine import CsmDims, CsmEngine, CsmWeights

    dims = CsmDims(**dataclasses.asdict(odims))
    w = CsmWeights(weights, dims)
    kv = torch.zeros(dims.num_hidden_layers, pages, 2, page, dims.num_key_value_heads, dims.head_dim, dtype=BF, device="cuda")
    return CsmEngine(w, kv, page, max_batch=max_batch, max_rows=max_rows)


def _i32(x):
    return torch.tensor(x, dtype=torch.int32, device="cuda")


def _check_step(got_logits, ref_logits, got_id, ref_id, st, tol=2e-2):
    ref = ref_logits.float()
    err = float((got_logits.float().cpu() - ref).abs().max() / ref.abs().max())
    st["max_err"] = max(st["max_err"], err)
    st["rows"] += 1
    if got_id != ref_id:
        top2 = torch.topk(ref, 2).values
        assert float(top2[0] - top2[1]) <= 2 * tol * float(ref.abs().max()), (got_id, ref_id, float(top2[0] - top2[1]))
        st["flips"] += 1


def test_csm_tiny_frames_against_reference_golden(golden_dir):
    from vox_serve_b200 import ops
    from vox_serve_b200.sampling import SamplingConfig

    gd = np.load(f"{golden_dir}/csm_tiny_frames.npz")
    odims = ocsm.CsmDims.tiny()
    weights = ocsm.synth_weights(odims, seed=int(gd["weight_seed"]))
    page = int(gd["page_size"])
    eng = _engine(odims, weights, page, pages=8, max_batch=2, max_rows=64)
    cfg = SamplingConfig(greedy=True)
    ids = torch.from_numpy(gd["prompt_ids"]).cuda()
    masks = torch.from_numpy(gd["prompt_masks"]).cuda()
    T0, N = ids.shape[0], odims.num_codebooks
    n_pages = (T0 + page - 1) // page
    ops.plan_rows(eng.bb.plan, _i32([0, T0]), _i32([0, n_pages]), _i32(list(range(n_pages))),
                  _i32([T0 - (n_pages - 1) * page]), 1, T0, page, eng.bb.chunk)
    st = dict(rows=0, flips=0, max_err=0.0)
    kv_len = T0
    for f in range(len(gd["frames"])):
        keep = []
        if f == 0:
            out = eng.prefill_frame(ids, masks, torch.arange(T0, dtype=torch.int32, device="cuda"), _i32([T0 - 1]),
                                    eng.bb.plan, cfg, keep_logits=keep)
        else:
            kv_len += 1
            n_pages = (kv_len + page - 1) // page
            ops.plan_rows(eng.bb.plan, None, _i32([0, n_pages]), _i32(list(range(n_pages))),
                          _i32([kv_len - (n_pages - 1) * page]), 1, 1, page, eng.bb.chunk)
            out = eng.decode_frame(1, _i32([kv_len - 1]), eng.bb.plan, cfg, keep_logits=keep)
        torch.cuda.synchronize()
        got = out[0].cpu().tolist()
        want = gd["frames"][f].tolist()
        assert got[N] == got[0]                                   # text column = codebook 0 (csm.py:693 repeat quirk)
        _check_step(keep[0][0], torch.from_numpy(gd["cb0_logits"][f]), got[0], want[0], st)
        for c in range(1, N):
            _check_step(keep[c][0], torch.from_numpy(gd["depth_logits"][f][c - 1]), got[c], want[c], st)
            if got[c] != want[c]:
                break                       # later codebooks of this frame were conditioned on a different id
        eng.frame[:N, 0] = torch.tensor(want, dtype=torch.int64, device="cuda")      # teacher forcing
    print("csm tiny vs reference golden:", st)
    assert st["max_err"] < 2e-2 and st["flips"] <= 2, st


def test_csm_batch_frames_one_graph_against_oracle():
    """Ragged batch of 3 requests, backbone head_dim 64 / depth head_dim 128 / GQA 4 like csm-1b, page 32; every decode
    frame is ONE CUDA-graph replay (backbone + 15 depth steps + samplers), compared request by request with the oracle."""
    from vox_serve_b200 import ops
    from vox_serve_b200.sampling import SamplingConfig

    odims = ocsm.CsmDims.tiny(hidden_size=512, num_hidden_layers=3, num_attention_heads=8, num_key_value_heads=2, head_dim=64,
                              intermediate_size=1024, num_codebooks=16, vocab_size=515, text_vocab_size=300,
                              depth_hidden_size=256, depth_num_hidden_layers=2, depth_num_attention_heads=4,
                              depth_num_key_value_heads=1, depth_head_dim=128, depth_intermediate_size=512)
    weights = ocsm.synth_weights(odims, seed=5)
    N, page, n_frames = odims.num_codebooks, 32, 4
    g = torch.Generator().manual_seed(3)
    lens = [37, 9, 64]
    prompts, masks = [], []
    for T in lens:
        ids = torch.randint(0, odims.vocab_size, (T, N + 1), generator=g)
        ids[:, -1] = torch.randint(0, odims.text_vocab_size, (T,), generator=g)
        m = torch.zeros(T, N + 1, dtype=torch.bool)
        n_text = T // 2
        m[:n_text, -1] = True                # text rows: only the text stream; audio rows: only the audio streams
        m[n_text:, :N] = True
        prompts.append(ids)
        masks.append(m)
    ref = [ocsm.generate_frames(weights, odims, p, m, n_frames, page_size=page) for p, m in zip(prompts, masks)]
    B = len(lens)
    pages_per = [(T + n_frames + page - 1) // page + 1 for T in lens]
    eng = _engine(odims, weights, page, pages=sum(pages_per), max_batch=4, max_rows=256)
    cfg = SamplingConfig(greedy=True)
    base = np.cumsum([0] + pages_per).tolist()
    kv_len = list(lens)

    def table():
        npg = [(kv + page - 1) // page for kv in kv_len]
        indptr = np.cumsum([0] + npg).tolist()
        indices = [base[r] + j for r in range(B) for j in range(npg[r])]
        last = [kv_len[r] - (npg[r] - 1) * page for r in range(B)]
        return _i32(indptr), _i32(indices), _i32(last)

    st = dict(rows=0, flips=0, max_err=0.0)
    qo = np.cumsum([0] + lens).tolist()
    indptr, indices, last = table()
    ops.plan_rows(eng.bb.plan, _i32(qo), indptr, indices, last, B, qo[-1], page, eng.bb.chunk)
    pos = torch.cat([torch.arange(T, dtype=torch.int32) for T in lens]).cuda()
    keep = []
    out = eng.prefill_frame(torch.cat(prompts).cuda(), torch.cat(masks).cuda(), pos, _i32([x - 1 for x in qo[1:]]), eng.bb.plan,
                            cfg, keep_logits=keep)
    graph = None
    d_pos = torch.zeros(B, dtype=torch.int32, device="cuda")
    d_indptr, d_indices, d_last = [torch.zeros_like(x) for x in table()]
    d_indices = torch.zeros(sum(pages_per), dtype=torch.int32, device="cuda")
    for f in range(n_frames):
        torch.cuda.synchronize()
        got = out[:B].cpu().tolist()
        for r in range(B):
            want = ref[r]["frames"][f]
            if keep:
                _check_step(keep[0][r], ref[r]["cb0_logits"][f], got[r][0], want[0], st)
                for c in range(1, N):
                    _check_step(keep[c][r], ref[r]["depth_logits"][f][c - 1], got[r][c], want[c], st)
                    if got[r][c] != want[c]:
                        break
            else:
                st["flips"] += int(got[r][:N] != want)
            eng.frame[:N, r] = torch.tensor(want, dtype=torch.int64, device="cuda")
        if f == n_frames - 1:
            break
        kv_len = [k + 1 for k in kv_len]
        indptr, indices, last = table()
        d_pos.copy_(_i32([k - 1 for k in kv_len]))
        d_indptr.copy_(indptr), d_last.copy_(last)
        d_indices[:indices.numel()].copy_(indices)
        if f == 0:
            # eager once (with logits kept), then capture: frames 2.. are single graph replays
            ops.plan_rows(eng.bb.plan, None, d_indptr, d_indices, d_last, B, B, page, eng.bb.chunk)
            keep = []
            out = eng.decode_frame(B, d_pos, eng.bb.plan, cfg, keep_logits=keep)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            n0 = ops.launch_count()
            with torch.cuda.stream(s):
                with torch.cuda.graph(graph, stream=s):
                    ops.plan_rows(eng.bb.plan, None, d_indptr, d_indices, d_last, B, B, page, eng.bb.chunk)
                    out_g = eng.decode_frame(B, d_pos, eng.bb.plan, cfg)
            torch.cuda.current_stream().wait_stream(s)
            st["graph_nodes"] = ops.launch_count() - n0
        else:
            keep = []
            graph.replay()
            out = out_g
    print("csm batch frames vs oracle:", st)
    assert st["max_err"] < 2e-2 and st["flips"] <= 3 and st["graph_nodes"] > 100, st


def test_csm_adapter_frame_device_equals_step_at_a_time_surface():
    """CSMModel (the reference adapter's contract, csm.py:315-789): the fused ``frame_device`` against the reference-style
    loop driven from outside -- ``forward`` -> ``sampling`` -> 2-row depth prefill -> ``depth_sampling`` -> 1-row depth
    decodes (cuda_graph_worker.py:1058-1160) -- through the FlashInfer-compatible wrappers; greedy ids must be equal."""
    from vox_serve_b200.flashinfer_utils import FlashInferDecodeWrapper, FlashInferPrefillWrapper
    from vox_serve_b200.model import load_model
    from vox_serve_b200.requests import Request
    from vox_serve_b200.sampling import SamplingConfig

    model = load_model("csm-synthetic-tiny:1", device="cuda:0", greedy=True)
    assert model.has_depth_transformer and model.n_codebooks == 9 and model.depth_n_codebooks == 8
    d, N, page, B = model.dims, model.dims.num_codebooks, 16, 2
    cfg = SamplingConfig(greedy=True)
    kw = dict(attn_buffer=None, n_qo_head=d.num_attention_heads, n_kv_head=d.num_key_value_heads,
              n_state=d.num_attention_heads * d.head_dim, page_size=page, device="cuda")
    g = torch.Generator().manual_seed(2)
    lens = [21, 7]
    ids = torch.randint(0, d.vocab_size, (sum(lens), N + 1), generator=g)
    ids[:, -1] = torch.randint(0, d.text_vocab_size, (sum(lens),), generator=g)
    masks = torch.ones(sum(lens), N + 1, dtype=torch.bool)
    masks[:, -1] = False
    masks[:5] = False
    masks[:5, -1] = True
    qo = [0, lens[0], sum(lens)]
    pages = [[0, 1], [2]]
    results = {}
    for mode in ("fused", "stepwise"):
        kv = torch.zeros(d.num_hidden_layers, 8, 2, page, d.num_key_value_heads, d.head_dim, dtype=BF, device="cuda")
        eng = model.engine_for(kv, page)
        pre = FlashInferPrefillWrapper(batch_size=B, max_seq_len=64, **kw)
        pre.plan(torch.tensor(qo, dtype=torch.int32), torch.tensor([0, 2, 3], dtype=torch.int32),
                 torch.tensor([0, 1, 2], dtype=torch.int32), torch.tensor([lens[0] - page, lens[1]], dtype=torch.int32))
        pos = torch.cat([torch.arange(n, dtype=torch.int32) for n in lens]).cuda()
        last = torch.tensor([q - 1 for q in qo[1:]], dtype=torch.int32, device="cuda")
        frames = []
        if mode == "fused":
            frames.append(model.frame_device(kv, pre, pos, B, ids.cuda(), masks.cuda(), last_rows=last, sampling_params=cfg).cpu())
        else:
            reqs = [Request(request_id=f"c{i}", prompt=None) for i in range(B)]
            for r, n in zip(reqs, lens):
                r.next_position_id = n + 1
            logits, hidden = model.forward(ids.cuda(), pos, pre, kv, input_masks=masks.cuda())
            out, x2 = model.sampling(logits[last.long()], hidden[last.long()], reqs, sampling_params=cfg)
            dpre = FlashInferPrefillWrapper(batch_size=B, max_seq_len=2 * B, n_qo_head=d.depth_num_attention_heads,
                                            n_kv_head=d.depth_num_key_value_heads, attn_buffer=None,
                                            n_state=d.depth_num_attention_heads * d.depth_head_dim, page_size=eng.depth_page,
                                            device="cuda")
            ddec = FlashInferDecodeWrapper(batch_size=B, n_qo_head=d.depth_num_attention_heads, attn_buffer=None,
                                           n_kv_head=d.depth_num_key_value_heads,
                                           n_state=d.depth_num_attention_heads * d.depth_head_dim, page_size=eng.depth_page,
                                           device="cuda")
            i32 = lambda x: torch.tensor(x, dtype=torch.int32)     # noqa: E731
            out = out.clone()
            for i in range(1, N):
                if i == 1:
                    dpre.plan(i32([0, 2, 4]), i32([0, 1, 2]), i32([0, 1]), i32([2, 2]))
                    lg = model.depth_forward(x2.view(2 * B, -1), torch.tensor([0, 1] * B, dtype=torch.int32, device="cuda"),
                                             dpre, eng.depth_kv)[1::2]
                else:
                    ddec.plan(i32([0, 1, 2]), i32([0, 1]), i32([i + 1, i + 1]))
                    lg = model.depth_forward(x, torch.full((B,), i, dtype=torch.int32, device="cuda"), ddec, eng.depth_kv)
                out[:, i], x = model.depth_sampling(lg, i, reqs, sampling_params=cfg)
            frames.append(out.cpu())
            assert [int(v) for v in reqs[0].lm_output_tokens[-1][0, :N]] == out[0, :N].tolist()
        results[mode] = frames
    assert torch.equal(results["fused"][0][:, :N], results["stepwise"][0][:, :N]), (results["fused"][0], results["stepwise"][0])
    assert torch.equal(results["fused"][0][:, N], results["fused"][0][:, 0])


def test_csm_postprocess_runs_the_mimi_decoder():
    """csm.py:771-785: [B, 10, 33] frame rows -> drop the text column, clamp to the Mimi tables, decode every chunk on
    its own -> [B, 1, 19200]; checked against the Mimi oracle on the adapter's own (seeded) decoder weights."""
    from oracle import mimi as omimi
    from vox_serve_b200.model.csm import CSMModel
    from vox_serve_b200.tokenizer.mimi import synthetic_state_dict

    model = CSMModel("csm-synthetic-tiny:4")
    mc = model.audio_decoder.cfg
    assert model.detokenize_interval == 10 and model.output_audio_length == 10 * mc.hop == 19200
    g = torch.Generator().manual_seed(0)
    N = model.dims.num_codebooks
    rows = torch.randint(0, model.dims.vocab_size, (3, 10, N + 1), generator=g)
    rows[0, 0, 0] = 5000                    # beyond the table: clamped to 2047 like the reference
    wav = model.postprocess(rows.cuda())
    assert wav.shape == (3, 1, 19200)
    ocfg = omimi.MimiConfig(**{f.name: getattr(mc, f.name) for f in __import__("dataclasses").fields(omimi.MimiConfig)})
    ref = omimi.decode(synthetic_state_dict(mc, 4), ocfg, rows[:, :, :-1].transpose(1, 2).clamp(0, 2047))
    assert float((wav.cpu() - ref).abs().max() / ref.abs().max()) < 2e-4
