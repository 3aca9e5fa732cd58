"""Why is the first msw_step_host call after a pause slower (0.25 vs 0.155 ms in bench.py's e2e series)?  Times the first
three calls after a synchronize + a busy-wait of X us on the host, for several X."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import minesweeper_ppo_b200 as m

N, K = 65536, 400
dev = torch.device("cuda")
cfg = m.EnvConfig(H=16, W=16, mine_count=40)
vec = m.VecMinesweeper(N, cfg, seed=0, api="torch")
log = torch.empty((K, N), dtype=torch.int32, device=dev)
vec.reset()
for t in range(K):
    vec.step_random(t, actions_out=log[t])
acts = log.cpu().pin_memory()
v = m.VecMinesweeper(N, cfg, seed=0, api="torch")
v.reset()
t = 0
for _ in range(20):
    v.step_host(acts[t], copy_obs=False, copy_infos=False); t += 1


def spin(us):
    t0 = time.perf_counter()
    while (time.perf_counter() - t0) * 1e6 < us:
        pass


for gap in (0, 20, 50, 100, 200, 500, 2000, 20000):
    rows = []
    for rep in range(5):
        torch.cuda.synchronize()
        spin(gap)
        ts = []
        for k in range(3):
            t0 = time.perf_counter()
            v.step_host(acts[t], copy_obs=False, copy_infos=False); t += 1
            ts.append((time.perf_counter() - t0) * 1e6)
        rows.append(ts)
    med = np.median(np.array(rows), axis=0)
    print(f"host idle {gap:6d} us before the call: 1st {med[0]:6.1f} us   2nd {med[1]:6.1f} us   3rd {med[2]:6.1f} us")
# the same with an event record in front of the first call (bench.py records e0 there)
for gap in (0, 200):
    rows = []
    for rep in range(5):
        torch.cuda.synchronize()
        spin(gap)
        e = torch.cuda.Event(enable_timing=True); e.record()
        ts = []
        for k in range(3):
            t0 = time.perf_counter()
            v.step_host(acts[t], copy_obs=False, copy_infos=False); t += 1
            ts.append((time.perf_counter() - t0) * 1e6)
        rows.append(ts)
    med = np.median(np.array(rows), axis=0)
    print(f"event record first, idle {gap:4d} us: 1st {med[0]:6.1f} us   2nd {med[1]:6.1f} us   3rd {med[2]:6.1f} us")
