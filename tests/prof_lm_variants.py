"""Dev tool (GPU): decode-step time of the CosyVoice2-0.5B and GLM-4-Voice-9B decoder stacks at their TRUE shapes (synthetic
weights) on the engine -- BASELINE.json configs[0] (single prompt, greedy) and the LM of configs[4] (batch 8).  One CUDA
graph per step: advance kv / positions on the device, plan, embed the previous ids, forward, greedy sample.
    python tests/prof_lm_variants.py cosyvoice2|glm [batch] [prompt_rows] [steps]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402
from vox_serve_b200.engine import BF16, hf_layer_names  # noqa: E402
from vox_serve_b200.lm_variants import CosyVoice2LM, GLMVoiceLM, cosyvoice2_dims, glm_voice_dims  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cosyvoice2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T0 = int(sys.argv[3]) if len(sys.argv) > 3 else 300
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
dev, page = "cuda", 128
g = torch.Generator(device=dev).manual_seed(0)


def rnd(*shape, std=0.02, mean=0.0):
    out = torch.empty(*shape, dtype=BF16, device=dev)
    flat = out.view(shape[0], -1) if len(shape) > 1 else out.view(-1, 1)
    step = max(1, (1 << 26) // max(1, flat.shape[1]))
    for r in range(0, flat.shape[0], step):
        flat[r:r + step].copy_(torch.randn(flat[r:r + step].shape, generator=g, dtype=torch.float32, device=dev) * std + mean)
    return out


t0 = time.perf_counter()
if which == "cosyvoice2":
    d = cosyvoice2_dims()
    P = CosyVoice2LM.PREFIX
    H, I, D = d.hidden_size, d.intermediate_size, d.head_dim
    hq, hkv = d.num_attention_heads * D, d.num_key_value_heads * D
    sd = {P + "embed_tokens.weight": rnd(151936, H, std=1.0), P + "norm.weight": rnd(H, std=0.1, mean=1.0),
          "llm_decoder.weight": rnd(d.vocab_size, H, std=0.16), "llm_decoder.bias": rnd(d.vocab_size, std=0.1),
          "speech_embedding.weight": rnd(d.vocab_size, H, std=1.0), "llm_embedding.weight": rnd(2, H, std=1.0)}
    for i in range(d.num_hidden_layers):
        n = hf_layer_names(i, P)
        sd[n["ln1"]], sd[n["ln2"]] = rnd(H, std=0.1, mean=1.0), rnd(H, std=0.1, mean=1.0)
        sd[n["q"]], sd[n["k"]], sd[n["v"]], sd[n["o"]] = rnd(hq, H), rnd(hkv, H), rnd(hkv, H), rnd(H, hq)
        sd[n["gate"]], sd[n["up"]], sd[n["down"]] = rnd(I, H), rnd(I, H), rnd(H, I)
        lp = f"{P}layers.{i}.self_attn."
        sd[lp + "q_proj.bias"], sd[lp + "k_proj.bias"], sd[lp + "v_proj.bias"] = rnd(hq, std=0.3), rnd(hkv, std=0.3), rnd(hkv, std=0.3)
else:
    d = glm_voice_dims()
    H, I, D = d.hidden_size, d.intermediate_size, d.head_dim
    qkv = H + 2 * D * d.num_key_value_heads
    sd = {"transformer.embedding.word_embeddings.weight": rnd(d.vocab_size, H, std=1.0),
          "transformer.encoder.final_layernorm.weight": rnd(H, std=0.1, mean=1.0),
          "transformer.output_layer.weight": rnd(d.vocab_size, H, std=0.16)}
    for i in range(d.num_hidden_layers):
        s = f"transformer.encoder.layers.{i}."
        sd[s + "input_layernorm.weight"], sd[s + "post_attention_layernorm.weight"] = rnd(H, std=0.1, mean=1.0), rnd(H, std=0.1, mean=1.0)
        sd[s + "self_attention.query_key_value.weight"], sd[s + "self_attention.query_key_value.bias"] = rnd(qkv, H), rnd(qkv, std=0.3)
        sd[s + "self_attention.dense.weight"] = rnd(H, H)
        sd[s + "mlp.dense_h_to_4h.weight"], sd[s + "mlp.dense_4h_to_h.weight"] = rnd(2 * I, H), rnd(H, I)
pages_req = (T0 + steps + 8 + page - 1) // page
kv = torch.zeros(d.num_hidden_layers, B * pages_req, 2, page, d.num_key_value_heads, d.head_dim, dtype=BF16, device=dev)
lm = (CosyVoice2LM if which == "cosyvoice2" else GLMVoiceLM)(sd, d, kv, page, max_rows=max(64, T0 + 8))
del sd
torch.cuda.empty_cache()
eng = lm.engine
setup_s = time.perf_counter() - t0
i32 = dict(dtype=torch.int32, device=dev)
indptr = torch.arange(B + 1, **i32) * pages_req
indices = torch.arange(B * pages_req, **i32)
# ---- prefill, one request at a time (T0 rows each) ----
t1 = time.perf_counter()
for r in range(B):
    npg = (T0 + page - 1) // page
    ops.plan_rows(eng.plan, torch.tensor([0, T0], **i32), torch.tensor([0, npg], **i32), indices[r * pages_req:r * pages_req + npg].contiguous(),
                  torch.tensor([T0 - (npg - 1) * page], **i32), 1, T0, page, eng.chunk)
    pos = torch.arange(T0, **i32)
    last = torch.tensor([T0 - 1], **i32)
    if which == "cosyvoice2":
        logits = lm.forward_embeds(rnd(T0, d.hidden_size, std=1.0), pos, last_rows=last)
    else:
        logits = lm.forward(torch.randint(0, d.vocab_size, (T0,), device=dev, dtype=torch.int32), pos, last_rows=last)
torch.cuda.synchronize()
prefill_ms = (time.perf_counter() - t1) * 1e3 / B
kv_len = torch.full((B,), T0, **i32)
pos = torch.full((B,), T0 - 1, **i32)
ids64 = torch.randint(0, min(d.vocab_size, 6000), (B,), device=dev)
ids32 = ids64.to(torch.int32)


def step():
    ops.decode_advance(kv_len, pos)
    ops.plan_rows(eng.plan, None, indptr, indices, None, B, B, page, eng.chunk, kv_len=kv_len)
    if which == "cosyvoice2":
        lg = lm.forward_speech_ids(ids32, pos)
    else:
        lg = lm.forward(ids32, pos)
    ops.sample(lg[:B], "greedy", out=ids64)
    ids32.copy_(ids64)


step()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
n0 = ops.launch_count()
with torch.cuda.stream(s):
    with torch.cuda.graph(gr, stream=s):
        step()
torch.cuda.current_stream().wait_stream(s)
nodes = ops.launch_count() - n0
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    gr.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
wbytes = lm.weights.streamed_bytes_per_step()
print(json.dumps({"model": which, "batch": B, "prompt_rows": T0, "steps": steps, "ms_per_step": ms, "tokens_per_s": B * 1e3 / ms,
                  "weights_gb_per_step": wbytes / 1e9, "weight_stream_tb_s": wbytes / (ms * 1e-3) / 1e12, "graph_nodes": nodes,
                  "prefill_ms_per_request_eager": prefill_ms, "final_kv_len": int(kv_len[0]), "setup_s": setup_s}))
