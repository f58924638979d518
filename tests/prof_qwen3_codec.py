"""Dev tool (GPU): one streaming decode_chunk of the Qwen3 12 Hz codec at the deployed widths (16 streams x 10 frames), for
ncu launch lists: python tests/prof_qwen3_codec.py [B] [iters]"""
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200.model.qwen3_tts import _synthetic_codec_state_dict  # noqa: E402
from vox_serve_b200.tokenizer.qwen3_codec import Qwen3CodecConfig, Qwen3TTSDecoder  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = Qwen3CodecConfig()
dec = Qwen3TTSDecoder(config=cfg, state_dict=_synthetic_codec_state_dict(cfg, 0))
codes = torch.randint(0, cfg.codebook_size, (B, cfg.num_quantizers, 10), device="cuda")
cache = dec.init_cache(B)
dec.decode_chunk(codes, cache)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()
e0.record()
for _ in range(iters):
    wav, _ = dec.decode_chunk(codes, cache)
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"qwen3 codec decode_chunk B={B} x 10 frames -> {tuple(wav.shape)}: {e0.elapsed_time(e1) / iters:.2f} ms")
