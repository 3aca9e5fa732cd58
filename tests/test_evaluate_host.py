"""Host-side statistics of minesweeper_ppo_b200.evaluate (restated from eval.py:54-90, 447-458), checked on CPU
against independent brute-force definitions.  (Importing the module needs torch, not a GPU.)"""
import math

import numpy as np


def test_auroc_is_the_pairwise_probability_without_ties():
    from minesweeper_ppo_b200.evaluate import _auroc
    rng = np.random.default_rng(0)
    scores = rng.permutation(400).astype(np.float64) / 400.0          # distinct scores: no tie handling involved
    labels = (rng.random(400) < 0.3).astype(np.float32)
    pos, neg = scores[labels == 1], scores[labels == 0]
    want = float((pos[:, None] > neg[None, :]).mean())
    assert abs(_auroc(labels, scores) - want) < 1e-12
    assert math.isnan(_auroc(np.ones(5, np.float32), rng.random(5)))   # one class only (eval.py:59-60)


def test_ece_bins_and_closed_last_bin():
    from minesweeper_ppo_b200.evaluate import _ece
    probs = np.array([0.0, 0.05, 0.5, 0.95, 1.0])                      # 1.0 belongs to the last (closed) bin
    labels = np.array([0.0, 1.0, 1.0, 1.0, 0.0])
    # bins of width 1/15: {0.0, 0.05} -> |0.5 - 0.025|, {0.5} -> |1 - 0.5|, {0.95, 1.0} -> |0.5 - 0.975|
    want = (2 / 5) * 0.475 + (1 / 5) * 0.5 + (2 / 5) * 0.475
    assert abs(_ece(probs, labels) - want) < 1e-12
    assert math.isnan(_ece(np.zeros(0), np.zeros(0)))


def test_wilson_interval():
    from minesweeper_ppo_b200.evaluate import _wilson
    lo, hi = _wilson(0, 96)
    assert lo == 0.0 and abs(hi - 0.0385) < 1e-4                       # the value the reference printed for 0 / 96
    lo, hi = _wilson(50, 100)
    assert abs((lo + hi) / 2 - 0.5) < 1e-12 and 0.40 < lo < 0.41 and 0.59 < hi < 0.60
    assert all(math.isnan(v) for v in _wilson(0, 0))
