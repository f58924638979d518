"""Timing probe (GPU): the GLM-4-Voice speech tokenizer at its true shapes (16 Whisper-large encoder layers of width
1280, 20 heads, 30 s of audio = 3000 mel frames -> 1500 positions -> 375 tokens), synthetic weights.  One JSON line.

    python tests/prof_glm_encoder.py > gpurun_out/glm_encoder.json
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import glm_encoder as oenc  # noqa: E402  (weight synthesis only)
from vox_serve_b200 import ops  # noqa: E402
from vox_serve_b200.encoder import GLMEncoderConfig, GLMWhisperVQEncoder  # noqa: E402


def main():
    d = oenc.GLMEncoderDims()
    sd = oenc.synth_state_dict(d, 0)
    cfg = GLMEncoderConfig(**{k: getattr(d, k) for k in d.__dataclass_fields__})
    enc = GLMWhisperVQEncoder(cfg, sd)
    del sd
    frames = 3000
    feats = torch.randn(1, d.num_mel_bins, frames, device="cuda").to(torch.bfloat16)
    mask = torch.ones(1, frames, dtype=torch.long, device="cuda")
    for _ in range(2):
        ids = enc(feats, mask)
    torch.cuda.synchronize()
    before = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 5
    for _ in range(n):
        ids = enc(feats, mask)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    params = sum(2 * 4 * d.d_model ** 2 + 2 * 2 * d.d_model * d.encoder_ffn_dim for _ in range(d.quantize_position))
    print(json.dumps({"workload": "GLM-4-Voice tokenizer, 30 s of audio (3000 mel frames -> 375 tokens), synthetic weights",
                      "ms_per_prompt_device": round(e0.elapsed_time(e1) / n, 3), "ms_per_prompt_wall": round(wall, 3),
                      "launches_per_prompt": (ops.launch_count() - before) // n, "layer_weight_bytes": params,
                      "tokens": int(ids.shape[1]), "distinct_ids": len(set(ids[0].tolist()))}))


if __name__ == "__main__":
    main()
