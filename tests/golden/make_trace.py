#!/usr/bin/env python
"""Record the at-scale reference traces (tests/golden/trace_*.npz) from the LIVE reference.

    python tests/golden/make_trace.py [A B C]

BASELINE.md section 4's C2 parity sub-run -- N=4,096 envs x 256 steps, 16x16x40 -- run on the
unmodified reference (baseline/_ref, installed by tools/install_reference.sh) with its own PCG64 mine
layouts; what is stored is a SHA-256 digest per step and per output field (obs, mask, reward, done,
infos, auxiliary maps, per-env state), see tests/trace.py.  About two minutes per series.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import trace as TR  # noqa: E402

N, T = 4096, 256

if __name__ == "__main__":
    for series in (sys.argv[1:] or ["A", "B", "C"]):
        cfg = TR.trace_cfg()
        t0 = time.time()
        fx = TR.record_reference_trace(series, N, T, cfg)
        np.savez_compressed(TR.fixture_path(series, N, T, cfg), **fx)
        print(f"series {series}: {N}x{T} steps in {time.time() - t0:.0f}s  wins={fx['wins']} losses={fx['losses']} "
              f"layouts={fx['layouts_placed']} noop={fx['noop_clicks']} max_new_reveals={fx['max_new_reveals']}", flush=True)
