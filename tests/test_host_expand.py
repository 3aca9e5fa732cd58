"""CPU: the host-side expansion of msw_step_host (msw_expand_obs_host: packed bitboards -> the reference's
fp32 observation planes / bool mask) against the oracle's encoder on mid-game states.  No GPU needed: the
function is pure host code inside the C-ABI library."""
import ctypes as C

import numpy as np
import pytest


def _pack(cells, HW):
    from minesweeper_ppo_b200.env import pack_boards
    return np.ascontiguousarray(pack_boards(cells, HW))


@pytest.mark.parametrize("H,W,M", [(16, 16, 40), (16, 30, 99), (30, 16, 99), (32, 32, 150), (8, 8, 10), (5, 7, 6),
                                   (1, 9, 2), (9, 1, 2), (3, 32, 20), (31, 31, 120)])
@pytest.mark.parametrize("threads", [1, 5])
def test_expand_matches_oracle_encoder(oracle, H, W, M, threads):
    from minesweeper_ppo_b200 import _lib
    L = _lib.load()
    O = oracle
    N, HW = 300, H * W
    cfg = O.OracleEnvConfig(H=H, W=W, mine_count=M, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    vec = O.OracleVecEnv(N, cfg, seed=3, nthreads=2)
    rng = np.random.default_rng(H * 100 + W)
    b = vec.reset()
    for t in range(6):                       # mid-game states: a mix of fresh, opened and auto-reset boards
        s = rng.random(b["action_mask"].shape)
        s[~b["action_mask"]] = -1
        b, _, _, _ = vec.step(s.argmax(1), tensor_infos=True)
        mines, rev = _pack(vec.mine.astype(bool), HW), _pack(vec.revealed.astype(bool), HW)
        meta = np.zeros((N, 4), np.int32)
        meta[:, 0] = vec.first_click_done
        obs = np.full((N, 10, H, W), np.nan, np.float32)
        mask = np.zeros((N, HW), bool)
        desc = _lib.EnvDesc(H, W, M, 1, 0.0, 0.0, 0.0, 0, 0, 0)
        rc = L.msw_expand_obs_host(C.byref(desc), mines.ctypes.data, rev.ctypes.data, meta.ctypes.data, N,
                                   obs.ctypes.data, mask.ctypes.data, threads)
        assert rc == 0
        assert np.array_equal(obs.view(np.uint32), b["obs"].view(np.uint32)), (H, W, t)
        assert np.array_equal(mask, b["action_mask"])


def test_expand_unaligned_and_partial_outputs(oracle):
    """obs at an address that is not 16-byte aligned (no streaming stores), obs-only and mask-only calls."""
    from minesweeper_ppo_b200 import _lib
    L = _lib.load()
    H = W = 16
    N, HW = 64, 256
    rng = np.random.default_rng(0)
    mine = rng.random((N, HW)) < 0.15
    rev = (rng.random((N, HW)) < 0.5) & ~mine
    meta = np.ones((N, 4), np.int32)
    meta[::3, 0] = 0                           # first_click_done = 0: count planes stay empty (env.py:181)
    raw = np.zeros(N * 10 * HW + 1, np.float32)
    obs = raw[1:].reshape(N, 10, H, W)         # 4-byte aligned only
    desc = _lib.EnvDesc(H, W, 40, 1, 0.0, 0.0, 0.0, 0, 0, 0)
    pm, pr = _pack(mine, HW), _pack(rev, HW)
    assert L.msw_expand_obs_host(C.byref(desc), pm.ctypes.data, pr.ctypes.data, meta.ctypes.data, N, obs.ctypes.data, None, 3) == 0
    counts = np.stack([np.asarray(oracle.adjacent_counts(m.reshape(H, W))) for m in mine]).reshape(N, HW)
    want = np.zeros((N, 10, HW), np.float32)
    want[:, 0] = rev
    for k in range(9):
        want[:, 1 + k] = rev & (counts == k) & (meta[:, 0:1] != 0)
    assert np.array_equal(obs.reshape(N, 10, HW), want)
    mask = np.zeros((N, HW), bool)
    assert L.msw_expand_obs_host(C.byref(desc), pm.ctypes.data, pr.ctypes.data, meta.ctypes.data, N, None, mask.ctypes.data, 0) == 0
    assert np.array_equal(mask, ~rev)
