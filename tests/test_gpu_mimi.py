"""Mimi decode on the GPU (CSM's vocoder, vox_serve/tokenizer/mimi.py:2993-3090) against the golden
file produced by the reference's own MimiModel.decode (tests/golden/mimi_tiny.npz, oracle/gen_golden.py:golden_mimi) and
against the oracle at the real widths.  fp32 end to end; the tolerance is on the error relative to the signal's scale,
because the GPU's tiled summation order differs from MKL's (the oracle itself is bit-exact against the reference)."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import mimi as omimi

pytestmark = pytest.mark.gpu

REL_TOL = 2e-4          # max |gpu - ref| / max |ref| for an fp32 pipeline of ~40 contractions (measured ~1e-5)


def _decoder(cfg, seed):
    from vox_serve_b200.tokenizer.mimi import MimiConfig, MimiDecoder

    sd = omimi.synth_state_dict(cfg, seed)
    return MimiDecoder(mimi_config=MimiConfig(**dataclasses.asdict(cfg)), state_dict=sd, num_codebooks=cfg.n_q), sd


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_mimi_decode_matches_reference_golden(golden_dir):
    gd = np.load(f"{golden_dir}/mimi_tiny.npz")
    cfg = omimi.MimiConfig.tiny()
    dec, _ = _decoder(cfg, int(gd["weight_seed"]))
    codes = torch.from_numpy(gd["codes"]).cuda()
    taps = {}
    wav = dec.decode(codes, taps)
    torch.cuda.synchronize()
    assert wav.shape == (3, 1, 5 * 1920) and wav.dtype == torch.float32
    for name, got in (("latent", taps["latent"]), ("transformer_out", taps["transformer_out"]), ("wav", wav)):
        ref = torch.from_numpy(gd[name])
        assert torch.isfinite(got).all()
        assert _rel(got.cpu(), ref) < REL_TOL, (name, _rel(got.cpu(), ref))
    # any int dtype is accepted, chunks are independent (no state carried between calls)
    again = dec.decode(codes.to(torch.int32))
    assert torch.equal(again, wav)
    one = dec.decode(codes[1:2])
    assert _rel(one.cpu(), torch.from_numpy(gd["wav"][1:2])) < REL_TOL


@pytest.mark.parametrize("B,K,T", [(1, 32, 10), (5, 32, 10), (2, 32, 3), (3, 8, 1)])
def test_mimi_decode_true_widths_matches_oracle(B, K, T):
    """The deployed widths (512-wide, 8 layers, 32 codebooks of 2048, ratios 8-6-5-4) on CSM's chunk shape
    (10 frames -> 19200 samples, csm.py:771-785), ragged batch / chunk lengths, and fewer codebooks than the quantizer
    holds (the reference's ``decode`` accepts any prefix of the codebooks, mimi.py:600-612)."""
    cfg = omimi.MimiConfig()
    dec, sd = _decoder(cfg, 5)
    g = torch.Generator().manual_seed(B * 100 + T)
    codes = torch.randint(0, cfg.bins, (B, K, T), generator=g)
    with torch.no_grad():
        ref = omimi.decode(sd, cfg, codes)
    wav = dec.decode(codes.cuda())
    assert wav.shape == ref.shape == (B, 1, T * 1920)
    assert _rel(wav.cpu(), ref) < REL_TOL, _rel(wav.cpu(), ref)


def test_mimi_rejects_what_it_cannot_decode():
    from vox_serve_b200._lib import VoxB200Error

    cfg = omimi.MimiConfig.tiny()
    dec, _ = _decoder(cfg, 1)
    with pytest.raises(VoxB200Error):
        dec.decode(torch.zeros(1, cfg.n_q, 4, dtype=torch.int64))               # CPU tensor: no CPU path
    with pytest.raises(VoxB200Error):
        dec.decode(torch.zeros(1, cfg.n_q + 1, 4, dtype=torch.int64, device="cuda"))
    with pytest.raises(VoxB200Error):
        dec.decode(torch.zeros(1, cfg.n_q, 40, dtype=torch.int64, device="cuda"))   # 80 positions > 64
    # out-of-table codes clamp instead of reading outside the codebook
    wav = dec.decode(torch.full((1, cfg.n_q, 2), 10 ** 6, dtype=torch.int64, device="cuda"))
    assert torch.isfinite(wav).all()
