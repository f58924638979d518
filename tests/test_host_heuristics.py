"""Host-side sizing rules (no GPU): split-K of the decode projections, gate/up tile height, and the receptive-field
walk of the SNAC decoder (vox_serve_b200/tokenizer/snac.py:_stage_ranges) against a brute-force dependency trace."""
import numpy as np
import pytest

from vox_serve_b200 import ops
from vox_serve_b200.tokenizer.snac import SNAC


def test_split_k_orpheus_decode_shapes():
    # (N, K) of the four projections at hidden 3072 / intermediate 8192, 32 rows, 148 SMs
    assert ops.choose_split_k(5120, 3072, 32, 148) == 4      # QKV: 40 tiles x 4
    assert ops.choose_split_k(3072, 3072, 32, 148) == 4      # O: 24 tiles x 4, 12 k-blocks per CTA
    assert ops.choose_split_k(3072, 8192, 32, 148) == 8      # down: 24 tiles x 8, 16 k-blocks per CTA
    assert ops.choose_split_k(156940, 3072, 32, 148) == 1    # lm_head: more tiles than SMs
    for N, K in [(256, 256), (1024, 1024), (3072, 64), (128, 4096)]:
        s = ops.choose_split_k(N, K, 8, 148)
        assert s >= 1 and (K + 63) // 64 >= s                # never more splits than k-blocks


def _brute_force_ranges(rates, t_latent, out_range):
    """needed[i] per stage computed by pushing index sets backwards through the layer graph"""
    lens = [t_latent]
    for s in rates:
        lens.append(lens[-1] * s)
    need = set(range(max(0, out_range[0] - 3), min(lens[-1], out_range[1] + 3)))      # final conv k7
    out = []
    for b in reversed(range(len(rates))):
        T, s = lens[b + 1], rates[b]
        units = []
        for dil in (9, 3, 1):
            units.append((min(need), max(need) + 1))
            need = {t + k * dil for t in need for k in range(-3, 4) if 0 <= t + k * dil < T}
        convtr = (min(need), max(need) + 1)
        pad = (s + 1) // 2
        src = set()
        for to in need:                      # to = ti * s - pad + k, k in [0, 2s)
            for k in range(2 * s):
                if (to + pad - k) % s == 0:
                    ti = (to + pad - k) // s
                    if 0 <= ti < lens[b]:
                        src.add(ti)
        need = src
        out.append((convtr, list(reversed(units))))
    return list(reversed(out)), (min(need), max(need) + 1)


@pytest.mark.parametrize("rates,t_latent,out_range", [((8, 8, 4, 2), 16, (2048, 4096)), ((8, 8, 4, 2), 16, (0, 8192)),
                                                      ((2, 2), 12, (5, 17)), ((4, 2), 8, (60, 64))])
def test_snac_stage_ranges_cover_the_receptive_field(rates, t_latent, out_range):
    m = SNAC.__new__(SNAC)
    m.decoder_rates = tuple(rates)
    blocks, first = m._stage_ranges(t_latent, out_range)
    ref_blocks, ref_first = _brute_force_ranges(rates, t_latent, out_range)
    # every computed range must contain what the brute-force trace needs (and stay inside the stage)
    for (c, units), (rc, runits) in zip(blocks, ref_blocks):
        assert c[0] <= rc[0] and c[1] >= rc[1]
        for u, ru in zip(units, runits):
            assert u[0] <= ru[0] and u[1] >= ru[1]
    assert first[0] <= ref_first[0] and first[1] >= ref_first[1]
    assert 0 <= first[0] < first[1] <= t_latent


def test_gate_up_tile_half_is_a_multiple_of_16():
    for inter in (8192, 2048, 4864, 11008):
        h = ops.gate_up_tile_half(inter, 148)
        assert h % 16 == 0 and 16 <= h <= 64
    assert ops.gate_up_tile_half(8192, 148) == 64


def test_tf32_hi_lo_split_is_fp32_grade():
    """The arithmetic identity behind vox_serve_b200/csrc/snac_mma.cu, emulated in numpy: x = hi + lo with hi = x with
    the 13 low mantissa bits cleared (exactly representable in tf32), lo = x - hi (exact in fp32; the tensor core then
    reads it with 10 mantissa bits).  lo_w*hi_x + hi_w*lo_x + hi_w*hi_x must reproduce the fp32 dot product to ~2^-20
    of the operand scale, where hi_w*hi_x alone (plain tf32) is ~2^-11."""
    rng = np.random.default_rng(0)

    def tf32(a):      # what kind::tf32 keeps of an fp32 operand (truncation; hi parts are unaffected by the mode)
        return (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

    K = 2048
    w = (rng.standard_normal((64, K)) / np.sqrt(K)).astype(np.float32)
    x = rng.standard_normal((K, 48)).astype(np.float32)
    w_hi, x_hi = tf32(w), tf32(x)
    w_lo, x_lo = tf32(w - w_hi), tf32(x - x_hi)
    assert np.array_equal(w_hi + (w - w_hi), w) and np.array_equal(x_hi + (x - x_hi), x)      # the split is exact
    exact = w.astype(np.float64) @ x.astype(np.float64)
    f64 = lambda a: a.astype(np.float64)      # noqa: E731
    split3 = f64(w_lo) @ f64(x_hi) + f64(w_hi) @ f64(x_lo) + f64(w_hi) @ f64(x_hi)
    plain = f64(w_hi) @ f64(x_hi)
    scale = np.sqrt((f64(w) ** 2).sum(1, keepdims=True) * (f64(x) ** 2).sum(0, keepdims=True))   # |w||x| per output
    err3 = np.abs(split3 - exact) / scale
    err1 = np.abs(plain - exact) / scale
    assert err3.max() < 2.0 ** -19, err3.max()
    assert err1.max() > 50 * err3.max()


def test_prefill_kernel_selection_rules():
    """Which attention kernel a plan runs on is a host-side decision on shapes only (ops.use_prefill_tiles /
    use_prefill_tc): stable under CUDA-graph capture, no device read."""
    from types import SimpleNamespace

    def plan(n_req, prefill=True):
        return SimpleNamespace(qo_indptr=object() if prefill else None, n_req=n_req)

    assert not ops.use_prefill_tiles(plan(32, prefill=False), 32, 128, 128)      # decode plan: one stream per row
    assert ops.use_prefill_tiles(plan(1), 133, 128, 128)                          # a lone prompt
    assert ops.use_prefill_tiles(plan(32), 164, 128, 128)                         # a prompt joining 31 decodes
    assert not ops.use_prefill_tiles(plan(64), 128, 128, 32)                      # 2-row depth-decoder prefill, batch 64
    assert not ops.use_prefill_tiles(plan(16), 32, 128, 32)
    assert not ops.use_prefill_tiles(plan(1), 133, 96, 128)                       # head_dim the tiled kernels do not take
    assert not ops.use_prefill_tc(plan(1), 435) and not ops.use_prefill_tc(plan(16), 800)
    assert not ops.use_prefill_tc(plan(32), 956)                                   # 7 prompts among 25 decode rows
    assert ops.use_prefill_tc(plan(4), 2400) and ops.use_prefill_tc(plan(1), 600)  # long prompts: the tcgen05 tiles
    assert ops.use_prefill_tc(plan(8), 1064)                                       # ... and big batches of prompts


def test_speech_tokenizer_config_and_mask_bounds():
    """GLMEncoderConfig.from_dict keeps the computation-shaping fields of the checkpoint's config.json and parks the rest;
    the per-row key bound equals the oracle's (which the golden test proves equal to the reference's additive mask)."""
    import torch

    from oracle import glm_encoder as oenc
    from vox_serve_b200.encoder.glm import GLMEncoderConfig, GLMWhisperVQEncoder

    cfg = GLMEncoderConfig.from_dict({"d_model": 1280, "encoder_attention_heads": 20, "quantize_position": 16,
                                      "vocab_size": 51866, "model_type": "whisper"})
    assert cfg.d_model == 1280 and cfg.quantize_position == 16 and cfg.extra == {"vocab_size": 51866, "model_type": "whisper"}
    for T, valid, block in ((80, 80, 16), (74, 66, 16), (1500, 1437, 200), (19, 17, 4)):
        am = torch.zeros(1, T, dtype=torch.long)
        am[:, :valid] = 1
        assert torch.equal(GLMWhisperVQEncoder.block_causal_bounds(valid, T, block, "cpu"),
                           oenc.block_causal_bounds(am, block))
