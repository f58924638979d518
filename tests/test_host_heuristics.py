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
