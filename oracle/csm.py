"""CPU restatement of the CSM backbone + depth-transformer frame step (BASELINE.json configs[3]: "CSM-1B
multi-codebook RVQ decode + Mimi codec"; SURVEY §8 row a24, the first "next" row of §8f).  TEST INFRASTRUCTURE ONLY:
imported by tests/ and oracle/gen_golden.py, never by the product path.

Follows ``vox_serve/model/csm.py`` and the worker's depth loop:
  * backbone = Llama-style decoder (``CsmAttention`` :55-115 -- no bias, llama-3.1 RoPE smoothing with the reference's
    defaults factor 32 / low 1 / high 4 / 8192 when ``rope_scaling`` carries none; ``CsmMLP`` :39-52; ``CsmDecoderLayer``
    :118-155; ``CsmBackboneModel`` :171-199), ``lm_head`` without bias for codebook 0;
  * frame input (``CSMModel.forward`` :637-663): the N audio ids of the previous frame embedded with per-codebook
    offsets (``CsmBackboneModelEmbeddings`` :158-168) plus the text id's embedding, masked and SUMMED over the N + 1
    columns;
  * depth decoder (``CsmDepthDecoderModel`` :202-232): ``inputs_embeds_projector`` (backbone width -> depth width),
    its own layers and norm, per-position head ``CsmCodebooksHead`` (:235-255): row at depth position p uses
    ``weight[p - 1]``;
  * the loop (``CSMModel.sampling`` :665-725, ``depth_sampling`` :749-769, ``CudaGraphWorker.run_lm_depth``
    cuda_graph_worker.py:1058-1160): codebook 0 from the backbone logits; depth step 1 is a 2-row prefill
    [backbone hidden state, embed(cb0)] at positions 0, 1 on a zeroed per-frame cache (one page per request); steps
    i = 2 .. N-1 are 1-row decodes at position i fed with embed(cb_{i-1} + (i-1) * vocab) from the BACKBONE's audio
    embedding table.
Pinned to the reference's own modules executed on CPU: tests/golden/csm_tiny_frames.npz
(oracle/gen_golden.py:golden_csm_frames).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from . import lm_ops

ROPE = dict(rope_scale=32.0, low_freq_factor=1.0, high_freq_factor=4.0, old_context_len=8192)   # csm.py:66-70 defaults


@dataclass
class CsmDims:
    # backbone (transformers.CsmConfig defaults: Llama-3.2-1B shape)
    hidden_size: int = 2048
    num_hidden_layers: int = 16
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: int = 64
    intermediate_size: int = 8192
    # codebooks
    num_codebooks: int = 32
    vocab_size: int = 2051
    text_vocab_size: int = 128256
    # depth decoder (CsmDepthDecoderConfig defaults)
    depth_hidden_size: int = 1024
    depth_num_hidden_layers: int = 4
    depth_num_attention_heads: int = 8
    depth_num_key_value_heads: int = 2
    depth_head_dim: int = 128
    depth_intermediate_size: int = 8192
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0

    @classmethod
    def tiny(cls, **kw):
        d = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                 intermediate_size=512, num_codebooks=8, vocab_size=64, text_vocab_size=100, depth_hidden_size=128,
                 depth_num_hidden_layers=2, depth_num_attention_heads=2, depth_num_key_value_heads=1,
                 depth_head_dim=64, depth_intermediate_size=256)
        d.update(kw)
        return cls(**d)


def _layer(prefix: str, i: int) -> Dict[str, str]:
    p = f"{prefix}layers.{i}."
    return {"ln1": p + "input_layernorm.weight", "ln2": p + "post_attention_layernorm.weight",
            "q": p + "self_attn.q_proj.weight", "k": p + "self_attn.k_proj.weight", "v": p + "self_attn.v_proj.weight",
            "o": p + "self_attn.o_proj.weight", "gate": p + "mlp.gate_proj.weight", "up": p + "mlp.up_proj.weight",
            "down": p + "mlp.down_proj.weight"}


BB, DD = "backbone_model.", "depth_decoder.model."


def _stack(w, prefix, n_layers, head_dim, eps, theta, h, position_ids, wrapper, kv_cache):
    t = h.shape[0]
    for i in range(n_layers):
        n = _layer(prefix, i)
        x = lm_ops.rms_norm(h, w[n["ln1"]], eps)
        q = F.linear(x, w[n["q"]]).view(t, -1, head_dim)
        k = F.linear(x, w[n["k"]]).view(t, -1, head_dim)
        v = F.linear(x, w[n["v"]]).view(t, -1, head_dim)
        q, k = lm_ops.apply_rope_pos_ids(q, k, position_ids, rope_theta=theta, **ROPE)
        wrapper.set_kv_cache(kv_cache[i], k, v)
        a = wrapper.run(q, kv_cache[i]).reshape(t, -1)
        h = h + F.linear(a, w[n["o"]])
        x = lm_ops.rms_norm(h, w[n["ln2"]], eps)
        h = h + F.linear(F.silu(F.linear(x, w[n["gate"]])) * F.linear(x, w[n["up"]]), w[n["down"]])
    return lm_ops.rms_norm(h, w[prefix + "norm.weight"], eps)


def frame_embeds(w, dims: CsmDims, input_ids: torch.Tensor, input_masks: torch.Tensor) -> torch.Tensor:
    """input_ids / input_masks [T, N + 1] (N audio columns, then the text column) -> [T, H] (csm.py:647-654)."""
    off = torch.arange(dims.num_codebooks) * dims.vocab_size
    audio = F.embedding(input_ids[:, :-1].long() + off, w[BB + "embed_tokens.embed_audio_tokens.weight"])
    text = F.embedding(input_ids[:, -1:].long(), w["embed_text_tokens.weight"])
    e = torch.cat([audio, text], dim=1)
    return (e * input_masks[:, :, None]).sum(dim=1)


def backbone_forward(w, dims: CsmDims, inputs_embeds, position_ids, wrapper, kv_cache) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (codebook-0 logits [T, vocab], last hidden [T, H])   (csm.py:290-300)"""
    h = _stack(w, BB, dims.num_hidden_layers, dims.head_dim, dims.rms_norm_eps, dims.rope_theta, inputs_embeds,
               position_ids, wrapper, kv_cache)
    return F.linear(h, w["lm_head.weight"]), h


def depth_forward(w, dims: CsmDims, inputs_embeds, position_ids, wrapper, kv_cache) -> torch.Tensor:
    """rows [R, H_backbone] at depth positions -> logits [R, vocab], row r through head ``weight[pos_r - 1]``
    (csm.py:213-232, 241-255, 302-312; position 0 wraps to the last head like the reference's negative index)"""
    h = F.linear(inputs_embeds, w[DD + "inputs_embeds_projector.weight"])
    h = _stack(w, DD, dims.depth_num_hidden_layers, dims.depth_head_dim, dims.rms_norm_eps, dims.rope_theta, h,
               position_ids, wrapper, kv_cache)
    head = w["depth_decoder.codebooks_head.weight"][position_ids.long() - 1]          # [R, H_depth, vocab]
    return torch.stack([F.linear(h[r], head[r].T) for r in range(h.shape[0])], dim=0)


def depth_loop_greedy(w, dims: CsmDims, hidden: torch.Tensor, cb0: int, page_size: int = 32, forced=None):
    """One request's codebooks 1 .. N-1 for one frame (greedy).  Returns (ids [N-1], logits [N-1, vocab]).  ``forced``
    (ids of codebooks 1 .. N-1): feed these back instead of the argmax (teacher forcing; the returned ids stay the
    oracle's own argmax of every step)."""
    N = dims.num_codebooks
    assert page_size >= N
    emb = w[BB + "embed_tokens.embed_audio_tokens.weight"]
    kv = torch.zeros(dims.depth_num_hidden_layers, 1, 2, page_size, dims.depth_num_key_value_heads,
                     dims.depth_head_dim, dtype=hidden.dtype)                      # zeroed per frame (:1076)
    x = torch.stack([hidden, emb[cb0 + 0 * dims.vocab_size]], dim=0)                # [2, H]: hidden state, embed(cb0)
    pre = lm_ops.PagedWrapperCPU("prefill", page_size)
    pre.plan([0, 2], [0, 1], [0], [2])
    logits = depth_forward(w, dims, x, torch.tensor([0, 1], dtype=torch.int32), pre, kv)[-1]
    ids, logs = [], []
    for i in range(1, N):
        logs.append(logits.float())
        tok = int(torch.argmax(logits.float()))
        ids.append(tok)
        if i == N - 1:
            break
        dec = lm_ops.PagedWrapperCPU("decode", page_size)
        dec.plan([0, 1], [0], [i + 2])                                              # kv length after this row
        fed = tok if forced is None else int(forced[i - 1])
        x = emb[fed + i * dims.vocab_size][None, :]
        logits = depth_forward(w, dims, x, torch.tensor([i + 1], dtype=torch.int32), dec, kv)[0]
    return ids, torch.stack(logs)


def synth_weights(dims: CsmDims, seed: int = 0, dtype=torch.bfloat16, head_scale: float = 8.0):
    """Seeded weights under the reference's state_dict names (csm.py:158-288)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    w = {}

    def stack(prefix, n_layers, H, I, nq, nkv, D):
        for i in range(n_layers):
            n = _layer(prefix, i)
            w[n["ln1"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
            w[n["ln2"]] = (1.0 + rnd(H, std=0.1).float()).to(dtype)
            w[n["q"]], w[n["k"]], w[n["v"]] = rnd(nq * D, H), rnd(nkv * D, H), rnd(nkv * D, H)
            w[n["o"]] = rnd(H, nq * D)
            w[n["gate"]], w[n["up"]], w[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
        w[prefix + "norm.weight"] = (1.0 + rnd(H, std=0.1).float()).to(dtype)

    H, Hd, N, V = dims.hidden_size, dims.depth_hidden_size, dims.num_codebooks, dims.vocab_size
    w[BB + "embed_tokens.embed_audio_tokens.weight"] = rnd(N * V, H, std=1.0)
    stack(BB, dims.num_hidden_layers, H, dims.intermediate_size, dims.num_attention_heads,
          dims.num_key_value_heads, dims.head_dim)
    w["lm_head.weight"] = rnd(V, H, std=0.02 * head_scale)
    w["embed_text_tokens.weight"] = rnd(dims.text_vocab_size, H, std=1.0)
    w[DD + "embed_tokens.weight"] = rnd(N * V, H, std=1.0)          # present in the checkpoint, unused by the path
    stack(DD, dims.depth_num_hidden_layers, Hd, dims.depth_intermediate_size, dims.depth_num_attention_heads,
          dims.depth_num_key_value_heads, dims.depth_head_dim)
    w[DD + "inputs_embeds_projector.weight"] = rnd(Hd, H, std=0.05)
    w["depth_decoder.codebooks_head.weight"] = rnd(N - 1, Hd, V, std=0.02 * head_scale)
    return w


def generate_frames(w, dims: CsmDims, prompt_ids: torch.Tensor, prompt_masks: torch.Tensor, n_frames: int,
                    page_size: int = 16) -> Dict[str, List]:
    """Single request: backbone prefill on the prompt rows, then ``n_frames`` frames -- codebook 0 greedy from the
    backbone, codebooks 1 .. N-1 from the depth loop, the finished frame fed back with the text column masked off
    (csm.py:665-725: ``input_masks[:, -1] = False``)."""
    T0, N = prompt_ids.shape[0], dims.num_codebooks
    n_pages = (T0 + n_frames + page_size - 1) // page_size + 1
    kv = torch.zeros(dims.num_hidden_layers, n_pages, 2, page_size, dims.num_key_value_heads, dims.head_dim,
                     dtype=w["lm_head.weight"].dtype)
    pages = list(range((T0 + page_size - 1) // page_size))
    pre = lm_ops.PagedWrapperCPU("prefill", page_size)
    pre.plan([0, T0], [0, len(pages)], pages, [T0 - (len(pages) - 1) * page_size])
    logits, hidden = backbone_forward(w, dims, frame_embeds(w, dims, prompt_ids, prompt_masks),
                                      torch.arange(T0, dtype=torch.int32), pre, kv)
    logits, hidden = logits[-1], hidden[-1]
    frames, cb0_logits, depth_logits, kv_len = [], [], [], T0
    mask = torch.ones(1, N + 1, dtype=torch.bool)
    mask[0, -1] = False
    for _ in range(n_frames):
        cb0_logits.append(logits.float())
        cb0 = int(torch.argmax(logits.float()))
        rest, dl = depth_loop_greedy(w, dims, hidden, cb0)
        frames.append([cb0] + rest)
        depth_logits.append(dl)
        kv_len += 1
        if (kv_len + page_size - 1) // page_size > len(pages):
            pages.append(len(pages))
        dec = lm_ops.PagedWrapperCPU("decode", page_size)
        dec.plan([0, len(pages)], pages, [kv_len - (len(pages) - 1) * page_size])
        row = torch.tensor([frames[-1] + [0]], dtype=torch.long)                 # text column: id 0, masked off
        lg, hd = backbone_forward(w, dims, frame_embeds(w, dims, row, mask),
                                  torch.tensor([kv_len - 1], dtype=torch.int32), dec, kv)
        logits, hidden = lg[0], hd[0]
    return {"frames": frames, "cb0_logits": cb0_logits, "depth_logits": depth_logits}
