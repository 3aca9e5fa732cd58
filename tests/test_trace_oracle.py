"""CPU: the oracle against the at-scale reference traces (tests/golden/trace_*.npz: 4,096 envs x 256
steps of the live reference, SHA-256 per step and field; tests/trace.py).  The reference's own mine
layouts are regenerated from NumPy's PCG64 streams (trace.LayoutReplayer)."""
import os

import pytest

import parity as P
import trace as TR

N, T = 4096, 256
# prefix lengths keep the CPU suite within minutes (hashing 50 MB per step); the -m gpu twin replays all 256
STEPS = {"A": 192, "B": 96, "C": 256}


@pytest.mark.parametrize("series", ["A", "B", "C"])
def test_oracle_matches_reference_trace(oracle, series):
    cfg = TR.trace_cfg()
    make = lambda c, n: P.OracleAdapter(oracle, c, n, nthreads=os.cpu_count() or 1)
    stats = TR.replay_trace(series, N, T, cfg, make, steps=STEPS[series])
    assert stats["layouts_placed"] > 0
    if series == "C":
        assert stats["wins"] > 5000          # the careful series exists to exercise wins at scale
