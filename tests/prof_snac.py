"""Dev tool (GPU, not a pytest file): time the SNAC decode of one batch of Orpheus windows (28 tokens -> 4 frames,
kept samples [2048, 4096)), the shape `run_detokenize` submits every 7th step.  `python tests/prof_snac.py [batch]`
prints one JSON line; under ncu it gives the per-kernel launch list of one decode."""
import json
import sys

import torch

sys.path.insert(0, ".")
from vox_serve_b200.tokenizer.snac import SNAC  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    dev = "cuda"
    m = SNAC(device=dev)
    m.load_state_dict(m.synthetic_state_dict(0))
    g = torch.Generator().manual_seed(1)
    codes = [torch.randint(0, 4096, (B, 4 * k), generator=g).to(dev, torch.int32) for k in (1, 2, 4)]
    noises = [torch.randn(s, generator=g).to(dev) for s in m.noise_shapes(B, 16)]
    for _ in range(3):
        wav = m.decode(codes, noises, out_range=(2048, 4096))
    torch.cuda.synchronize()
    # one decode as a CUDA graph (how the resident loop replays it): GPU time without the host's launch cost
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            wav = m.decode(codes, noises, out_range=(2048, 4096))
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    if len(sys.argv) > 3 and sys.argv[3] == "trace":
        # in-situ timeline of the tensor-core launches of one replay (block 0 of each: %globaltimer marks)
        from vox_serve_b200 import _lib
        cap = 1024
        buf = torch.zeros(4 + 4 * cap, dtype=torch.int64, device=dev)
        buf[1] = cap
        _lib.check(_lib.load().vb_set_trace(buf.data_ptr()), "vb_set_trace")
        g.replay()
        torch.cuda.synchronize()
        _lib.check(_lib.load().vb_set_trace(None), "vb_set_trace")
        h = buf.cpu()
        n = min(int(h[0]), cap)
        recs = sorted(h[4:4 + 4 * n].view(n, 4).tolist())
        names = {60: "tc kernel", 61: " setup done", 62: " first stage ready", 63: " all MMAs issued", 64: " accumulator done"}
        t0 = recs[0][0]
        for a, b, kid, aux in recs:
            print(f"{(a - t0) / 1e3:9.2f} us  {names.get(kid, kid):22s} k-blocks {aux:3d}  dur {(b - a) / 1e3:7.2f} us")
    print(json.dumps({"snac_decode_ms": round(e0.elapsed_time(e1) / reps, 4), "batch": B,
                      "checksum": float(wav.double().abs().sum())}))


if __name__ == "__main__":
    main()
