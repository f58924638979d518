"""Metric definitions of benchmarks/openloop.py against the reference client's (benchmark/goodput.py:186-215, 354-363)."""
import numpy as np

from benchmarks.openloop import arrivals, viability


def test_poisson_arrivals_rate_and_seed():
    a = arrivals(8.0, 60.0, seed=42)
    assert a == arrivals(8.0, 60.0, seed=42) and a != arrivals(8.0, 60.0, seed=43)
    assert all(x < y for x, y in zip(a, a[1:])) and a[-1] < 60.0
    assert abs(len(a) / 60.0 - 8.0) < 1.0                      # ~480 arrivals, sd ~22
    gaps = np.diff(np.array([0.0] + a))
    assert abs(gaps.mean() - 1 / 8.0) < 0.02 and abs(gaps.std() - 1 / 8.0) < 0.03      # exponential: mean = sd


def test_streaming_viability_definition():
    # chunk i is on time iff the audio of chunks 0..i-1 outlasts its arrival latency since chunk 0
    dur = [0.0853] * 4
    assert viability([0.0, 0.05, 0.10, 0.15], dur) == (100.0, True)
    v = viability([0.0, 0.05, 0.20, 0.21], dur)                                    # chunk 2 late: 0.1706 < 0.20
    assert abs(v[0] - 200.0 / 3) < 1e-9 and v[1] is False
    assert viability([0.0], [0.0853]) is None
