"""`analyze_avoidability` with the reference's call shape and result type.

Mirrors `minesweeper/avoidability.py:145-394` for a `vec.envs[i]` view.  The whole batch is analysed by
one kernel launch (msw_avoidability: frontier components, unit + subset rules, exact per-component
search) the first time any env of a step is asked for, and cached until the env steps again -- the
per-env Python loop of eval.py:381-398 then only reads arrays.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Set

import numpy as np


@dataclass
class AvoidabilityResult:                     # avoidability.py:19-29
    avoidable: bool
    forced_safe_cells: Set[int]
    component_sizes: List[int]
    chosen_is_forced_safe: bool
    chosen_component_size: Optional[int]

    @property
    def count_forced_safe_cells(self) -> int:
        return len(self.forced_safe_cells)


def analyze_avoidability(state, chosen_cell: Optional[int], *, component_threshold: int = 22) -> AvoidabilityResult:
    """`component_threshold` is accepted for signature compatibility: both branches of the reference
    (:366-373) compute the same set, and so does the kernel."""
    vec, i = state._vec, state._i
    arr = vec.avoidability()
    flags = int(arr["flags"][i])
    if flags & 8:
        raise RuntimeError(f"analyze_avoidability: exact search of env {i} exceeded its step budget; "
                           "call vec.avoidability(search_budget=...) with a larger budget first")
    safe = {int(k) for k in np.flatnonzero(arr["safe"][i])}
    sizes = [int(s) for s in arr["comp_size"][i] if s > 0]
    chosen_safe, chosen_size = False, None
    if chosen_cell is not None and flags & 4:
        cell = int(chosen_cell)
        if flags & 2:                          # :239-249, :340-341, :379-380
            lab = int(arr["comp_of_cell"][i][cell])
            if lab >= 0:
                chosen_size = int(arr["comp_size"][i][lab])
                chosen_safe = cell in safe
        elif not state.revealed.flat[cell]:    # :176-179: no frontier at all
            chosen_size = 1
    return AvoidabilityResult(bool(flags & 1), safe, sizes, chosen_safe, chosen_size)
