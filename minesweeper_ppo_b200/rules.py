"""Eval-side analytics with the reference's call shape.

`analyze_forced_modules(state)` mirrors `minesweeper/rules.py:206-259` for a `vec.envs[i]` view: the
pairwise subset rule on ground-truth mines ("subset_reveal").  The whole batch is analysed by one
kernel launch (msw_forced_subset) the first time any env of a step is asked for, and cached until
the env steps again -- the per-env Python loop of eval.py:350-398 then only reads sets.
"""
from __future__ import annotations

from typing import Dict, Set

import numpy as np


def analyze_forced_modules(state) -> Dict[str, Set[int]]:
    vec, i = state._vec, state._i
    row = vec.forced_subset()[i]
    return {"subset_reveal": {int(k) for k in np.flatnonzero(row)}}
