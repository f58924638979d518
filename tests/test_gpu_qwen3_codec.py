"""Qwen3-TTS 12 Hz codec decoder, streaming form, on the GPU (vox_serve/tokenizer/qwen3_codec.py:
1541-1667) against the golden file produced by the reference's own Qwen3TTSTokenizerV2Decoder.forward_chunk on CPU
(tests/golden/qwen3_codec_tiny.npz: three consecutive chunks of 5, 5, 3 frames with every cache carried over) and against the
oracle at a wider configuration with the deployed head geometry.  fp32; tolerance on the error relative to the signal scale."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import qwen3_codec as oq

pytestmark = pytest.mark.gpu
REL_TOL = 5e-4        # fp32 pipeline of ~90 contractions with sin / exp activations (fp32 SIMT kernels: measured ~1e-5)
REL_TOL_TC = 1e-3     # layers >= 64 channels wide run on the tcgen05 tf32 hi/lo kernel (~2e-5 per layer; measured 8e-4 after
#                       ~60 such layers with SnakeBeta between them): the waveform contract of north_star is 1e-3


def _decoder(cfg, seed):
    from vox_serve_b200.tokenizer.qwen3_codec import Qwen3CodecConfig, Qwen3TTSDecoder

    sd = oq.synth_state_dict(cfg, seed)
    return Qwen3TTSDecoder(config=Qwen3CodecConfig(**dataclasses.asdict(cfg)), state_dict=sd), sd


def _rel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-6))


def test_qwen3_codec_three_chunks_match_reference_golden(golden_dir):
    gd = np.load(f"{golden_dir}/qwen3_codec_tiny.npz")
    cfg = oq.Qwen3CodecConfig.tiny()
    dec, _ = _decoder(cfg, int(gd["weight_seed"]))
    cache = dec.init_cache(2)
    for i in range(3):
        codes = torch.from_numpy(gd[f"codes{i}"]).cuda()
        wav, cache2 = dec.decode_chunk(codes, cache)
        assert cache2 is cache and wav.shape == (2, 1, codes.shape[2] * cfg.hop)
        assert _rel(wav.cpu(), torch.from_numpy(gd[f"wav{i}"])) < REL_TOL, (i, _rel(wav.cpu(), torch.from_numpy(gd[f"wav{i}"])))
    torch.cuda.synchronize()
    # the state after three chunks: window contents (zeros where nothing was written yet), offsets, every conv cache
    assert cache.position_offset.tolist() == gd["position_offset"].tolist() == [13, 13]
    assert _rel(cache.attention_cache.cpu(), torch.from_numpy(gd["attention_cache"])) < REL_TOL
    assert _rel(cache.pre_conv_cache.cpu(), torch.from_numpy(gd["pre_conv_cache"])) < REL_TOL
    for name in ("upsample_conv_caches", "decoder_conv_caches", "transconv_caches"):
        for j, t in enumerate(getattr(cache, name)):
            assert _rel(t.cpu(), torch.from_numpy(gd[f"{name}.{j}"])) < REL_TOL, (name, j)


def test_qwen3_codec_streams_are_independent_and_caches_stack():
    """Batch items do not interact, and per-request caches stack / split through DecoderCache like the reference's
    (tokenizer/base.py:8-173): decoding two streams together equals decoding each with its own cache."""
    from vox_serve_b200.tokenizer.base import DecoderCache

    cfg = oq.Qwen3CodecConfig.tiny()
    dec, _ = _decoder(cfg, 3)
    g = torch.Generator().manual_seed(1)
    chunks = [torch.randint(0, cfg.codebook_size, (2, cfg.num_quantizers, 4), generator=g).cuda() for _ in range(3)]
    both = dec.init_cache(2)
    singles = [dec.init_cache(1), dec.init_cache(1)]
    for c in chunks:
        wav, _ = dec.decode_chunk(c, both)
        for r in range(2):
            w1, _ = dec.decode_chunk(c[r:r + 1], singles[r])
            assert torch.equal(w1, wav[r:r + 1])
    stacked = type(both).cat(singles)                 # what the worker does with per-request caches before a batched call
    assert isinstance(stacked, DecoderCache) and torch.equal(stacked.attention_cache, both.attention_cache)
    assert torch.equal(stacked.decoder_conv_caches[5], both.decoder_conv_caches[5])
    assert torch.equal(stacked[1:2].transconv_caches[2], singles[1].transconv_caches[2])


@pytest.mark.parametrize("B,T", [(1, 10), (3, 10), (2, 1)])
def test_qwen3_codec_deployed_head_geometry_matches_oracle(B, T):
    """16 heads of 64 (no GQA), 72-slot window, rates 8-5-4-3 after 2 x 2 upsampling, 16 codebooks, CSM-sized chunk of 10
    frames; widths reduced (decoder 256 instead of 1536) so that the CPU oracle finishes in seconds.  Two chunks."""
    cfg = oq.Qwen3CodecConfig(latent_dim=256, codebook_dim=128, codebook_size=256, decoder_dim=256, hidden_size=1024 // 4,
                              intermediate_size=512, head_dim=64, num_attention_heads=4, num_hidden_layers=3, num_key_value_heads=4)
    assert cfg.hop == 1920 and cfg.sliding_window == 72
    dec, sd = _decoder(cfg, 9)
    g = torch.Generator().manual_seed(B * 10 + T)
    cache, ocache = dec.init_cache(B), oq.init_cache(cfg, B)
    for _ in range(2):
        codes = torch.randint(0, cfg.codebook_size, (B, cfg.num_quantizers, T), generator=g)
        ref, ocache = oq.forward_chunk(sd, cfg, codes, ocache)
        wav, cache = dec.decode_chunk(codes.cuda(), cache)
        assert wav.shape == ref.shape == (B, 1, T * 1920)
        assert dec.tc, "this configuration must exercise the tensor-core convolutions"
        assert _rel(wav.cpu(), ref) < REL_TOL_TC, _rel(wav.cpu(), ref)


def test_qwen3_codec_rejects_what_it_cannot_decode():
    from vox_serve_b200._lib import VoxB200Error

    cfg = oq.Qwen3CodecConfig.tiny()
    dec, _ = _decoder(cfg, 1)
    with pytest.raises(ValueError):
        dec.decode_chunk(torch.zeros(1, cfg.num_quantizers + 1, 4, dtype=torch.int64, device="cuda"))
    with pytest.raises(VoxB200Error):
        dec.decode_chunk(torch.zeros(1, cfg.num_quantizers, cfg.sliding_window, dtype=torch.int64, device="cuda"))
    with pytest.raises(VoxB200Error):
        dec.decode_chunk(torch.zeros(1, cfg.num_quantizers, 4, dtype=torch.int64))
