"""Host-side logic of the Mimi decoder that needs no GPU: state-dict loading (codebook materialisation, the per-phase
re-packing of the transposed-conv weights) checked against torch's own conv_transpose1d, and the synthetic generator's
key set against the reference checkpoint's decode-side names (as restated in oracle/mimi.py)."""
import dataclasses

import torch
import torch.nn.functional as F

from oracle import mimi as omimi
from vox_serve_b200.tokenizer.mimi import MimiConfig, MimiDecoder, seanet_layout, synthetic_state_dict


def test_load_packs_transposed_convs_per_output_phase():
    cfg = omimi.MimiConfig.tiny()
    sd = omimi.synth_state_dict(cfg, 17)
    dec = MimiDecoder(mimi_config=MimiConfig(**dataclasses.asdict(cfg)), state_dict=sd, device="cpu")
    g = torch.Generator().manual_seed(0)
    for kind, idx in seanet_layout(dec.cfg):
        if kind != "convtr":
            continue
        W = sd[f"decoder.model.{idx}.convtr.convtr.weight"]
        cin, cout, k = W.shape
        s, T = k // 2, 5
        x = torch.randn(2, cin, T, generator=g)
        ref = F.conv_transpose1d(x, W, stride=s)[..., : T * s]          # causal: the rightmost K - S outputs are trimmed
        taps = torch.cat([x, F.pad(x, (1, 0))[..., :-1]], 1)             # [B, 2 Cin, T]: x[n], x[n - 1]
        got = torch.einsum("rok,bkt->botr", dec.w[f"d{idx}.w"], taps).reshape(2, cout, T * s)
        assert torch.allclose(got, ref, atol=1e-5), idx
    # codebooks: embedding_sum / clamp(usage, eps), rvq_first then rvq_rest
    p = "quantizer.rvq_rest.vq.layers.2._codebook."
    want = sd[p + "embedding_sum"] / sd[p + "cluster_usage"].clamp(min=cfg.codebook_eps)[:, None]
    assert torch.equal(dec.w["codebooks"][3], want) and dec.w["codebooks"].shape == (cfg.n_q, cfg.bins, cfg.codebook_dim)


def test_synthetic_generator_covers_exactly_the_decode_side_keys():
    ours = synthetic_state_dict(MimiConfig(), 0)
    ref = {k: v.shape for k, v in omimi.synth_state_dict(omimi.MimiConfig(), 0).items()
           if "input_proj" not in k and "_initialized" not in k}
    assert {k: v.shape for k, v in ours.items()} == ref
    assert MimiConfig().hop == 1920 and MimiConfig().sample_rate == 24000
