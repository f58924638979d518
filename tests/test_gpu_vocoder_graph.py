"""Device noise generator (vb_randn, the NoiseBlock input of vox_serve/tokenizer/snac.py:206-212) and the CUDA-graphed
per-step vocoder of ModelWorker.run_detokenize (cuda_graph_worker.py:1162-1280 replays a detokenizer graph too)."""
import numpy as np
import pytest
import torch

from oracle import orpheus as oorph, snac as osnac

pytestmark = pytest.mark.gpu


def test_randn_statistics_determinism_and_state_advance():
    from vox_serve_b200 import ops

    n = 1 << 22
    a = ops.randn(n, seed=123, offset=7)
    b = ops.randn(n, seed=123, offset=7)
    c = ops.randn(n, seed=123, offset=8)
    assert torch.equal(a, b) and not torch.equal(a, c)
    x = a.double()
    assert abs(x.mean().item()) < 3e-3 and abs(x.var().item() - 1.0) < 5e-3
    assert abs((x ** 3).mean().item()) < 1e-2 and abs((x ** 4).mean().item() - 3.0) < 3e-2     # skewness, kurtosis
    assert x.abs().max().item() > 4.5 and torch.isfinite(a).all()
    # neighbouring outputs (same Philox block, both Box-Muller branches) and distant ones are uncorrelated
    for lag in (1, 2, 3, 4, 1000):
        assert abs((x[:-lag] * x[lag:]).mean().item()) < 3e-3, lag
    # two different streams are uncorrelated
    assert abs((x * c.double()).mean().item()) < 3e-3
    # tail sizes that are not a multiple of 4, and the device-side state: offset advances once per call
    st = torch.tensor([123, 7, 0], dtype=torch.int64, device="cuda")
    d = ops.randn(1001, rng_state=st)
    assert torch.equal(d, a[:1001]) and st.tolist() == [123, 8, 0]
    e = ops.randn(1001, rng_state=st)
    assert torch.equal(e, c[:1001]) and st.tolist() == [123, 9, 0]
    # a Kolmogorov-Smirnov check against the normal CDF
    s = torch.sort(x[: 1 << 18]).values.cpu().numpy()
    from scipy.stats import norm
    ks = np.abs(norm.cdf(s) - (np.arange(len(s)) + 0.5) / len(s)).max()
    assert ks < 4e-3, ks


def test_vocoder_graph_replay_equals_eager_launch_sequence():
    """run_detokenize's graph (window gather -> de-interleave -> SNAC -> PCM16 with device-drawn noise) must deliver
    the bytes of the eager launch sequence started from the same noise-stream offset, for several batch sizes, and
    fresh noise on every replay."""
    from tests.e2e_harness import build_models
    from vox_serve_b200 import ops

    dims = oorph.OrpheusDims.tiny(vocab_size=156940, stop_token_id=128258, audio_id_base=128266)
    dims.max_tokens = 200
    worker, _ = build_models(dims, osnac.SnacConfig.tiny(), 3, 8, 16, 128)
    dec = worker.model.audio_decoder
    assert dec.noise_source is None and worker._vocoder_graphable()
    W = worker.detokenize_interval
    g = torch.Generator().manual_seed(1)
    hist = torch.randint(dims.audio_id_base, dims.vocab_size, worker.history.shape, generator=g, dtype=torch.int32)
    worker.history.copy_(hist)
    st = dec.ensure_noise_state()
    for n in (1, 3, 8):
        wh = worker.win_host.numpy()
        for i in range(n):
            wh[0, i], wh[1, i], wh[2, i] = i, 7 * i, W if i % 2 == 0 else W - 5       # slot, first, n_valid
        worker.win_dev.copy_(worker.win_host)
        st[1] = 100 + n
        st[2] = 0
        worker._vocoder_body(n)
        torch.cuda.synchronize()
        eager = worker.voc_pcm[:n].clone()
        gph, nodes = worker._capture_vocoder(n)
        assert nodes > 10
        worker.voc_pcm.zero_()
        st[1] = 100 + n
        gph.replay()
        torch.cuda.synchronize()
        assert torch.equal(worker.voc_pcm[:n], eager), n
        assert int(st[1]) == 100 + n + 1
        gph.replay()
        torch.cuda.synchronize()
        assert not torch.equal(worker.voc_pcm[:n], eager), "a replay must draw fresh NoiseBlock noise"
        assert worker.voc_pcm[:n].abs().max().item() > 0
