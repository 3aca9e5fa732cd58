"""Not a test (not collected): host throughput of the reference-shaped result expansion at C2 size, FULL against
DELTA, on a played trajectory produced by the oracle (checker-side use only).

    python tests/host_expand_speed.py [envs] [threads]
"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minesweeper_ppo_b200 import _lib                      # noqa: E402
from minesweeper_ppo_b200.env import pack_boards           # noqa: E402
from oracle import oracle as O                             # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
TH = int(sys.argv[2]) if len(sys.argv) > 2 else 0
H = W = 16
HW = 256
L = _lib.load()
cfg = O.OracleEnvConfig(H=H, W=W, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
vec = O.OracleVecEnv(N, cfg, seed=0, nthreads=os.cpu_count() or 1)
rng = np.random.default_rng(1)
b = vec.reset()
states = []
for t in range(16):
    s = rng.random(b["action_mask"].shape, dtype=np.float32)
    s[~b["action_mask"]] = -1
    b, _, _, _ = vec.step(s.argmax(1), tensor_infos=True)
    if t >= 6:
        meta = np.zeros((N, 4), np.int32)
        meta[:, 0] = vec.first_click_done
        states.append((np.ascontiguousarray(pack_boards(vec.mine.astype(bool), HW)),
                       np.ascontiguousarray(pack_boards(vec.revealed.astype(bool), HW)), meta))
del b
desc = _lib.EnvDesc(H, W, 40, 1, 0, 0, 0, 0, 0, 0)
SW = L.msw_shadow_words(H, W)


def aligned(n_floats):
    raw = np.empty(n_floats + 32, np.float32)
    off = ((-raw.ctypes.data) % 64) // 4
    return raw[off:off + n_floats], raw


for sets in (0, 1, 2):
    pool = []
    for s in range(max(1, sets)):
        o, keep = aligned(N * 10 * HW)
        pool.append([o.reshape(N, 10, H, W), np.empty((N, HW), bool), np.zeros((N, SW), np.uint64), 0, keep])
    times = []
    for t, (pm, pr, meta) in enumerate(states):
        e = pool[t % len(pool)]
        t0 = time.perf_counter()
        if sets == 0:
            L.msw_expand_obs_host(C.byref(desc), pm.ctypes.data, pr.ctypes.data, meta.ctypes.data, N, e[0].ctypes.data,
                                  e[1].ctypes.data, TH)
        else:
            L.msw_expand_obs_host_delta(C.byref(desc), pm.ctypes.data, pr.ctypes.data, meta.ctypes.data, N, e[0].ctypes.data,
                                        e[1].ctypes.data, e[2].ctypes.data, e[3], TH)
            e[3] = 1
        times.append(time.perf_counter() - t0)
    steady = times[max(1, sets) + 1:]
    name = "full expansion" if sets == 0 else f"delta, {sets} result array set(s)"
    print(f"{name:32s}: first call {times[0] * 1e3:7.2f} ms, steady {np.median(steady) * 1e3:7.2f} ms (min {min(steady) * 1e3:.2f})  "
          f"{N / np.median(steady) / 1e6:7.2f} M env/s   (threads={TH or os.cpu_count()})")
    del pool
