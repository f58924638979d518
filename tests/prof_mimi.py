"""Dev tool (GPU): Mimi decode of B ten-frame chunks at the deployed widths (CSM's vocoder call, csm.py:771-785), timed
with CUDA events eagerly and as one CUDA-graph replay; also the target of the ncu capture in tests/gpu_round2.sh.
    python tests/prof_mimi.py [B] [iters]"""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200 import ops  # noqa: E402
from vox_serve_b200.tokenizer.mimi import MimiConfig, MimiDecoder, synthetic_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = MimiConfig()
dec = MimiDecoder(mimi_config=cfg, state_dict=synthetic_state_dict(cfg, 0))
codes = torch.randint(0, cfg.bins, (B, cfg.n_q, 10), device="cuda")
for _ in range(2):
    wav = dec.decode(codes)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = ops.launch_count()
e0.record()
for _ in range(iters):
    wav = dec.decode(codes)
e1.record()
torch.cuda.synchronize()
launches = (ops.launch_count() - n0) // iters
eager = e0.elapsed_time(e1) / iters
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        wav = dec.decode(codes)
torch.cuda.current_stream().wait_stream(s)
g.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    g.replay()
e1.record()
torch.cuda.synchronize()
graph = e0.elapsed_time(e1) / iters
audio_s = B * 10 * cfg.hop / cfg.sample_rate
print(f"mimi decode B={B} x 10 frames -> {tuple(wav.shape)}: {launches} launches, eager {eager:.3f} ms, graph {graph:.3f} ms "
      f"({audio_s / (graph / 1e3):.0f} audio-s/s for the vocoder alone)")
