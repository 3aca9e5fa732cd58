"""Host throughput of msw_expand_obs_host (the format conversion inside msw_step_host) at C2 size."""
import ctypes as C, numpy as np, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minesweeper_ppo_b200 import _lib
L = _lib.load()
N, H, W = 65536, 16, 16; HW = 256
rng = np.random.default_rng(0)
pm = rng.integers(0, 2**31, size=(N, 8), dtype=np.int32) & rng.integers(0, 2**31, size=(N, 8), dtype=np.int32) & rng.integers(0, 2**31, size=(N, 8), dtype=np.int32)
pr = rng.integers(0, 2**31, size=(N, 8), dtype=np.int32)
meta = np.ones((N, 4), np.int32)
raw = np.empty((N * 10 * HW + 32,), np.float32)
off = (-raw.ctypes.data % 64) // 4
obs = raw[off:off + N * 10 * HW].reshape(N, 10, H, W); obs_s = raw[off + 4:off + 4 + N * 10 * HW].reshape(N, 10, H, W); obs_u = raw[off + 1:off + 1 + N * 10 * HW].reshape(N, 10, H, W); mask = np.empty((N, HW), bool)
desc = _lib.EnvDesc(H, W, 40, 1, 0, 0, 0, 0, 0, 0)
obs[:] = 0
for name, o in (("streaming stores (16B aligned)", obs_s), ("regular stores (unaligned obs)", obs_u)):
    for th in (1, 4, 0):
        best = 1e9
        for rep in range(3):
            t = time.perf_counter()
            L.msw_expand_obs_host(C.byref(desc), pm.ctypes.data, pr.ctypes.data, meta.ctypes.data, N, o.ctypes.data, mask.ctypes.data, th)
            best = min(best, time.perf_counter() - t)
        print(f"{name:32s} threads={th:2d}: {best * 1e3:7.2f} ms  {N / best / 1e6:6.2f} M env/s  {N * 10496 / best / 1e9:6.1f} GB/s written")
