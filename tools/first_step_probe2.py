"""Reproduces bench.py's e2e timed loop (warm-up, rows, events, barrier, loop with the done-flag read) and toggles its
pieces to find what makes the FIRST timed call slow (0.25-0.36 ms against 0.155)."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import minesweeper_ppo_b200 as m

N, K = 65536, 600
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = m.EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
vec = m.VecMinesweeper(N, cfg, seed=0, api="torch")
log = torch.empty((K, N), dtype=torch.int32, device=dev)
vec.reset()
for t in range(K):
    vec.step_random(t, actions_out=log[t])
acts = log.cpu().pin_memory()
del vec, log
v = m.VecMinesweeper(N, cfg, seed=0, api="torch")
v.reset()
pos = 0


def leg(events=True, record_after_t0=True, collect=False, rows_first=True, steps=12, warm=5, read=True):
    global pos
    pin = None
    for _ in range(warm):
        pin = v.step_host(acts[pos], copy_obs=False, copy_infos=False); pos += 1
    rows = [acts[pos + i] for i in range(steps)] if rows_first else None
    if events:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e1.record()
    if collect:
        gc.collect()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if events and record_after_t0:
        e0.record()
    stamps = [t0]
    done = 0
    for i in range(steps):
        a = rows[i] if rows_first else acts[pos + i]
        pin = v.step_host(a, copy_obs=False, copy_infos=False)
        if read:
            done += int(np.count_nonzero(pin["done"].numpy()))
        stamps.append(time.perf_counter())
    if events:
        e1.record()
    torch.cuda.synchronize()
    pos += steps
    return [round(1e3 * (b - a), 3) for a, b in zip(stamps[:-1], stamps[1:])]


for name, kw in (("as bench.py", {}), ("as bench.py (again)", {}), ("no event record after t0", dict(record_after_t0=False)),
                 ("no events at all", dict(events=False)), ("with gc.collect()", dict(collect=True)),
                 ("rows sliced inside the loop", dict(rows_first=False)), ("no done-flag read", dict(read=False)),
                 ("no events, no read", dict(events=False, read=False))):
    s = leg(**kw)
    print(f"{name:32s}: first {s[0]:.3f}  then {' '.join('%.3f' % x for x in s[1:8])}")
