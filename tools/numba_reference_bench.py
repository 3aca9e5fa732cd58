#!/usr/bin/env python
"""Time the REAL reference env (Python + numba, baseline/_ref) on this box's host cores.

BASELINE.md section 4, "CPU baseline, timed in the same run on the same box":
  (i)  reference as shipped: one process / one core, `VecMinesweeper(N).step` for N in {64, 1024},
       16x16x40 and 16x30x99, numba warmed (JIT / cache load excluded), timing around `vec.step` only,
       actions `s=rng.random(mask.shape); s[~mask]=-1; s.argmax(1)` with `np.random.default_rng(1)`
       (the loop of scripts/profile_env.py:17-31 with the action sampler moved out of the timed region);
  (ii) fan-out: P = len(os.sched_getaffinity(0)) worker processes (harness-level: the reference has no
       multiprocessing), each with its own `VecMinesweeper(N/P)`; aggregate env-steps/s over the slowest
       worker's wall time, P and the CPU model reported.

Prints ONE JSON object.  Run as a subprocess by bench.py (a fresh interpreter: no CUDA context is
forked).  The reference is imported unmodified; nothing of this repo's package is on the timed path.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _import_reference(ref_dir: str):
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "msw_numba_cache_bench"))
    os.environ.setdefault("NUMBA_NUM_THREADS", "1")
    sys.dont_write_bytecode = True
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import minesweeper.env as E
    from minesweeper.env_numba import HAS_ENV_NUMBA
    assert HAS_ENV_NUMBA, "numba flood fill not active"
    assert os.path.realpath(E.__file__).startswith(os.path.realpath(ref_dir))
    return E


def time_steps(E, H: int, W: int, mines: int, n_envs: int, steps: int, warmup: int, seed: int = 0,
               barrier=None):
    """env-steps/s of vec.step alone (actions drawn outside the timed region)."""
    import numpy as np
    cfg = E.EnvConfig(H=H, W=W, mine_count=mines, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    vec = E.VecMinesweeper(n_envs, cfg, seed=seed)
    rng = np.random.default_rng(1 + seed)
    mask = vec.reset()["action_mask"]
    total = 0.0
    for t in range(warmup + steps):
        if t == warmup and barrier is not None:
            barrier.wait()
        s = rng.random(mask.shape)
        s[~mask] = -1.0
        a = s.argmax(1).astype(np.int32)
        t0 = time.perf_counter()
        batch, _, _, _ = vec.step(a)
        dt = time.perf_counter() - t0
        mask = batch["action_mask"]
        if t >= warmup:
            total += dt
    return n_envs * steps / total, total


def _worker(ref_dir, H, W, mines, n_envs, steps, warmup, idx, barrier, q):
    try:
        try:
            os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[idx % len(os.sched_getaffinity(0))]})
        except Exception:
            pass
        E = _import_reference(ref_dir)
        t_wall0 = None
        rate, total = time_steps(E, H, W, mines, n_envs, steps, warmup, seed=idx, barrier=barrier)
        q.put((idx, rate, total, None))
    except Exception as e:  # pragma: no cover
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((idx, 0.0, 0.0, repr(e)))


def fanout(ref_dir, H, W, mines, procs, envs_per_proc, steps, warmup):
    ctx = mp.get_context("spawn")
    barrier, q = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(ref_dir, H, W, mines, envs_per_proc, steps, warmup, i, barrier, q))
          for i in range(procs)]
    for p in ps:
        p.start()
    res = [q.get(timeout=600) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    errs = [r[3] for r in res if r[3]]
    if errs:
        return {"error": errs[0]}
    slowest = max(r[2] for r in res)
    return {"processes": procs, "envs_per_process": envs_per_proc, "steps": steps,
            "env_steps_per_s": procs * envs_per_proc * steps / slowest,
            "sum_of_worker_rates": sum(r[1] for r in res), "slowest_worker_s": slowest}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-dir", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--quick", action="store_true", help="a few seconds in total (tests)")
    ap.add_argument("--no-fanout", action="store_true")
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(a.ref_dir, "minesweeper", "env.py")):
        print(json.dumps({"unavailable": f"{a.ref_dir} is missing (tools/install_reference.sh)"}))
        return
    E = _import_reference(a.ref_dir)
    t0 = time.time()
    time_steps(E, 8, 8, 10, 8, 3, 1)                     # numba JIT / cache load, excluded from every number
    jit_s = time.time() - t0
    k = 0.2 if a.quick else 1.0
    out = {"impl": "yakvrz/minesweeper-ppo minesweeper.env.VecMinesweeper (unmodified, numba flood fill)",
           "cpu_model": cpu_model(), "host_threads": len(os.sched_getaffinity(0)), "numba_warmup_s": round(jit_s, 2),
           "timing": "perf_counter around vec.step only; actions drawn outside the timed region"}
    r, s = time_steps(E, 16, 16, 40, 64, int(300 * k), 5)
    out["one_core_N64"] = {"env_steps_per_s": r, "board": "16x16x40", "steps": int(300 * k), "seconds": s}
    r, s = time_steps(E, 16, 16, 40, 1024, max(3, int(24 * k)), 2)
    out["one_core_N1024"] = {"env_steps_per_s": r, "board": "16x16x40", "steps": max(3, int(24 * k)), "seconds": s}
    r, s = time_steps(E, 16, 30, 99, 1024, max(3, int(16 * k)), 2)
    out["one_core_N1024_expert"] = {"env_steps_per_s": r, "board": "16x30x99", "steps": max(3, int(16 * k)), "seconds": s}
    if not a.no_fanout:
        P = len(os.sched_getaffinity(0))
        out["fanout_P"] = fanout(a.ref_dir, 16, 16, 40, P, 256, max(4, int(200 * k)), 2)      # >= 1 s per worker
        out["fanout_P"]["board"] = "16x16x40"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
