"""DecoderCache (vox_serve/tokenizer/base.py:8-173): slice / copy_from / cat / to over nested codec state, checked
against the reference's own class when the pip-installed copy (baseline/_ref) is present."""
from dataclasses import dataclass
from typing import Any

import pytest
import torch

from vox_serve_b200.tokenizer.base import DecoderCache


def _make(base):
    @dataclass
    class Inner(base):
        conv: torch.Tensor = None
        steps: Any = None

    @dataclass
    class Outer(base):
        kv: Any = None
        tail: Any = None
        inner: Any = None
        offset: Any = None
        name: str = "c"

    def build(b, seed):
        g = torch.Generator().manual_seed(seed)
        return Outer(kv=[torch.randn(b, 2, 3, generator=g), (torch.randn(b, 4, generator=g), None)],
                     tail={"a": torch.randn(b, 5, generator=g)}, inner=Inner(conv=torch.randn(b, 6, generator=g), steps=3),
                     offset=torch.arange(b), name="c")

    return build


def _flat(c):
    return [c.kv[0], c.kv[1][0], c.tail["a"], c.inner.conv, c.offset]


def test_decoder_cache_slice_cat_copy_to():
    build = _make(DecoderCache)
    a, b = build(2, 0), build(3, 1)
    ab = type(a).cat([a, b])
    assert [t.shape[0] for t in _flat(ab)] == [5] * 5 and ab.inner.steps == 3 and ab.name == "c" and ab.kv[1][1] is None
    back = ab[2:]
    assert all(torch.equal(x, y) for x, y in zip(_flat(back), _flat(b)))
    idx = ab[torch.tensor([0, 4])]
    assert torch.equal(idx.inner.conv, torch.stack((a.inner.conv[0], b.inner.conv[2])))
    dst = build(3, 7)
    dst.copy_from(b)
    assert all(torch.equal(x, y) for x, y in zip(_flat(dst), _flat(b))) and dst.kv[0] is not b.kv[0]
    assert all(t.device.type == "cpu" for t in _flat(ab.to("cpu")))
    with pytest.raises(TypeError):
        dst.copy_from(a.inner)
    with pytest.raises(ValueError):
        type(a).cat([])
    a.inner.steps = 4
    with pytest.raises(TypeError):
        type(a).cat([a, b])            # plain members must agree


def test_decoder_cache_matches_reference_class():
    from tests import ref_dropin

    if not ref_dropin.available():
        pytest.skip("baseline/_ref (pip-installed reference) is absent")
    import sys

    if ref_dropin.REF not in sys.path:
        sys.path.insert(0, ref_dropin.REF)
    from vox_serve.tokenizer.base import DecoderCache as RefCache

    ours, theirs = _make(DecoderCache), _make(RefCache)
    for op in ("cat", "slice", "index", "copy"):
        res = []
        for build in (ours, theirs):
            a, b = build(2, 0), build(3, 1)
            if op == "cat":
                r = type(a).cat([a, b])
            elif op == "slice":
                r = type(a).cat([a, b])[1:4]
            elif op == "index":
                r = type(a).cat([a, b])[torch.tensor([4, 0, 2])]
            else:
                r = build(3, 5)
                r.copy_from(b)
            res.append(_flat(r) + [r.inner.steps, r.name])
        for x, y in zip(*res):
            assert torch.equal(x, y) if torch.is_tensor(x) else x == y


def test_index_copy_and_zero_rows_round_trip():
    from vox_serve_b200.tokenizer.qwen3_codec import Qwen3TTSDecoderCache

    def mk(B, v):
        return Qwen3TTSDecoderCache(attention_cache=torch.full((B, 2, 3), v), position_offset=torch.full((B,), int(v), dtype=torch.long),
                                    pre_conv_cache=torch.full((B, 4), v), upsample_conv_caches=[torch.full((B, 2), v)],
                                    decoder_conv_caches=[torch.full((B, 1), v), torch.full((B, 5), v)], transconv_caches=[])

    slots = mk(4, 0.0)
    part = mk(2, 7.0)
    idx = torch.tensor([3, 1])
    slots.index_copy_(idx, part)
    assert slots.attention_cache[:, 0, 0].tolist() == [0.0, 7.0, 0.0, 7.0] and slots.position_offset.tolist() == [0, 7, 0, 7]
    back = slots[idx]
    assert torch.equal(back.decoder_conv_caches[1], part.decoder_conv_caches[1])
    slots.zero_rows_(torch.tensor([1]))
    assert slots.pre_conv_cache[:, 0].tolist() == [0.0, 0.0, 0.0, 7.0]
    slots.zero_rows_(3)
    assert float(slots.upsample_conv_caches[0].abs().sum()) == 0.0
