"""Dev tool (GPU): BASELINE.json configs[3] shape -- CSM-1B (synthetic weights at the transformers.CsmConfig default
shapes), batch 64, Mimi vocoder every 10 frames -- served through the worker API by the in-process scheduler.
Prints frames/s and audio-s/s (one frame = 80 ms of audio) of the steady state, plus the decode-frame graph alone.
    python tests/prof_csm.py [batch] [prompt_rows] [frames]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from vox_serve_b200.model.csm import CSMModel  # noqa: E402
from vox_serve_b200.requests import Request  # noqa: E402
from vox_serve_b200.scheduler import Scheduler  # noqa: E402
from vox_serve_b200.worker import CudaGraphWorker  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T0 = int(sys.argv[2]) if len(sys.argv) > 2 else 600
F = int(sys.argv[3]) if len(sys.argv) > 3 else 60
t0 = time.perf_counter()
model = CSMModel("csm-synthetic:0", max_tokens=T0 + 400)
model.stop_token_id = -1                  # fixed-length streams: the stop frame (codebook 0 == 0) never ends a request
page = 128
pages = B * ((T0 + 400 + page - 1) // page + 1)
worker = CudaGraphWorker("csm-synthetic:0", max_batch_size=B, max_num_pages=pages, page_size=page, model=model,
                         max_prefill_tokens=1024)
import os  # noqa: E402
setup_s = time.perf_counter() - t0
N = model.dims.num_codebooks
g = torch.Generator().manual_seed(0)
sched = Scheduler(worker)
for i in range(B):
    ids = torch.randint(1, model.dims.vocab_size, (T0, N + 1), generator=g)
    ids[:, -1] = torch.randint(0, 1000, (T0,), generator=g)
    m = torch.zeros(T0, N + 1, dtype=torch.bool)
    m[: T0 // 4, -1] = True
    m[T0 // 4:, :N] = True
    sched.submit(Request(request_id=f"c{i}", prompt=(ids, m)))
t1 = time.perf_counter()
state = sched.run_async(B + 2)            # one prefill per step: everybody past prefill
torch.cuda.synchronize()
prefill_s = time.perf_counter() - t1
worker.capture_decode_graphs([B])
state = sched.run_async(12, state)        # warm: graphs, vocoder batch sizes
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
# the decode-frame graph alone (the staging buffer still describes the last step: all B requests decoding)
gr = worker.decode_graphs[B]
assert len(state[1]) == B
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    gr.replay()
e1.record()
torch.cuda.synchronize()
print(json.dumps({"workload": f"CSM-1B synthetic, batch {B}, {T0}-row prompts, Mimi every 10 frames", "frames": F,
                  "ms_per_frame_step": ms / F, "frames_per_s": B * F / (ms / 1e3), "audio_sec_per_sec": audio / (ms / 1e3),
                  "audio_sec_per_sec_from_frames": B * F * 0.08 / (ms / 1e3), "decode_frame_graph_ms": e0.elapsed_time(e1) / 10,
                  "graph_nodes": worker._graph_nodes.get(B), "setup_s": setup_s, "prefill_phase_s": prefill_s}))
if state[0] is not None:
    state[0].close()          # the carried request-state coroutine of the last step
