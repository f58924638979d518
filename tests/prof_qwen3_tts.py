"""Dev tool (GPU): BASELINE.json configs[2] shape per replica -- Qwen3-TTS-1.7B (synthetic weights at the in-tree default shapes:
talker 28 L x 2048, code predictor 5 L x 1024, 16 code groups), continuous batch of B streams, streaming 12 Hz codec every 10
frames -- through scheduler + DepthModelWorker.  Prints frames/s and audio-s/s (one frame = 80 ms) of the steady state, the
decode-frame graph alone and one codec chunk call alone.
    python tests/prof_qwen3_tts.py [batch] [prompt_rows] [frames]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from vox_serve_b200.model.qwen3_tts import Qwen3TTSModel  # noqa: E402
from vox_serve_b200.requests import Request  # noqa: E402
from vox_serve_b200.scheduler import Scheduler  # noqa: E402
from vox_serve_b200.worker import CudaGraphWorker  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 50
F = int(sys.argv[3]) if len(sys.argv) > 3 else 60
t0 = time.perf_counter()
model = Qwen3TTSModel("qwen3-tts-synthetic:0", max_tokens=T0 + 400, stop_token_id=-1)
page = 128
pages = B * ((T0 + 400 + page - 1) // page + 1)
worker = CudaGraphWorker("qwen3-tts-synthetic:0", max_batch_size=B, max_num_pages=pages, page_size=page, model=model,
                         max_prefill_tokens=1024)
setup_s = time.perf_counter() - t0
d = model.dims
N = d.num_code_groups
g = torch.Generator().manual_seed(0)
sched = Scheduler(worker)
for i in range(B):
    ids = torch.zeros(T0, N + 1, dtype=torch.int64)
    ids[:, -1] = torch.randint(0, 1000, (T0,), generator=g)
    ids[:, 0] = torch.randint(0, 2048, (T0,), generator=g)
    m = torch.zeros(T0, N + 1, dtype=torch.bool)
    m[T0 - 8:, -1] = True
    feats = (torch.randn(T0, d.hidden_size, generator=g) * 0.1).to(torch.bfloat16)
    sched.submit(Request(request_id=f"q{i}", prompt=(ids, m, feats)))
t1 = time.perf_counter()
state = sched.run_async(B + 2)
torch.cuda.synchronize()
prefill_s = time.perf_counter() - t1
worker.capture_decode_graphs([B])
state = sched.run_async(12, state)
torch.cuda.synchronize()
a0 = sched.audio_seconds()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t2 = time.perf_counter()
e0.record()
state = sched.run_async(F, state)
e1.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t2
ms = max(e0.elapsed_time(e1), wall * 1e3)
audio = sched.audio_seconds() - a0
gr = worker.decode_graphs[B]
assert len(state[1]) == B
e0.record()
for _ in range(10):
    gr.replay()
e1.record()
torch.cuda.synchronize()
graph_ms = e0.elapsed_time(e1) / 10
codes = torch.randint(0, 2048, (B, N, 10), device="cuda")
cache = model.audio_decoder.init_cache(B)
model.audio_decoder.decode_chunk(codes, cache)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    model.audio_decoder.decode_chunk(codes, cache)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"workload": f"Qwen3-TTS-1.7B synthetic, batch {B}, {T0}-row prompts, codec every 10 frames", "frames": F,
                  "ms_per_frame_step": ms / F, "frames_per_s": B * F / (ms / 1e3), "audio_sec_per_sec": audio / (ms / 1e3),
                  "decode_frame_graph_ms": graph_ms, "graph_nodes": worker._graph_nodes[B],
                  "codec_chunk_ms_for_batch": e0.elapsed_time(e1) / 3, "setup_s": setup_s, "prefill_phase_s": prefill_s}))
if state[0] is not None:
    state[0].close()
