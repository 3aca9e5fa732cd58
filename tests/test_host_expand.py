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


@pytest.mark.parametrize("H,W,M", [(16, 16, 40), (16, 30, 99), (30, 16, 99), (8, 8, 10), (5, 7, 6), (1, 9, 2), (3, 32, 20),
                                   (31, 31, 120)])
@pytest.mark.parametrize("sets,threads,align", [(1, 1, 64), (2, 5, 64), (2, 3, 4)])
def test_delta_expansion_tracks_a_trajectory(oracle, H, W, M, sets, threads, align):
    """msw_expand_obs_host_delta over a played trajectory (opened boards, no-op clicks, auto-resets): `sets` result
    array pairs are used round-robin as VecMinesweeper's pool does (so a pair is `sets` steps stale when it is
    reused), each with its own shadow; after every call the arrays equal the oracle's encoder bit for bit."""
    from minesweeper_ppo_b200 import _lib
    L = _lib.load()
    O = oracle
    N, HW = 257, H * W
    SW = L.msw_shadow_words(H, W)
    assert SW == (10 * HW + 63) // 64 + 1
    cfg = O.OracleEnvConfig(H=H, W=W, mine_count=M, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    vec = O.OracleVecEnv(N, cfg, seed=5, nthreads=2)
    rng = np.random.default_rng(H * 1000 + W + sets)
    desc = _lib.EnvDesc(H, W, M, 1, 0.0, 0.0, 0.0, 0, 0, 0)
    pool = []
    for s in range(sets):
        raw = np.full(N * 10 * HW + 32, np.nan, np.float32)                       # garbage before the first call
        off = ((-raw.ctypes.data) % 64) // 4 + (0 if align == 64 else 1)          # 64-byte aligned or only 4-byte aligned
        obs = raw[off:off + N * 10 * HW].reshape(N, 10, H, W)
        mask = np.ones((N, HW), bool) if s else np.zeros((N, HW), bool)
        pool.append([obs, mask, np.full((N, SW), 0x5555555555555555, np.uint64), 0, raw])
    b = vec.reset()
    for t in range(14):
        if t % 5 == 4:                                                         # any cell: includes no-op clicks
            a = rng.integers(0, HW, size=N)
        else:
            s_ = rng.random(b["action_mask"].shape)
            s_[~b["action_mask"]] = -1
            a = s_.argmax(1)
        b, _, _, _ = vec.step(a, tensor_infos=True)
        mines, rev = _pack(vec.mine.astype(bool), HW), _pack(vec.revealed.astype(bool), HW)
        meta = np.zeros((N, 4), np.int32)
        meta[:, 0] = vec.first_click_done
        e = pool[t % sets]
        # h_meta may be NULL (odd steps): first_click_done is implied by revealed != 0, which is how msw_step_host calls it
        rc = L.msw_expand_obs_host_delta(C.byref(desc), mines.ctypes.data, rev.ctypes.data, meta.ctypes.data if t % 2 == 0 else None, N,
                                         e[0].ctypes.data, e[1].ctypes.data, e[2].ctypes.data, e[3], threads)
        assert rc == 0
        e[3] = 1
        assert np.array_equal(e[0].view(np.uint32), b["obs"].view(np.uint32)), (H, W, t)
        assert np.array_equal(e[1], b["action_mask"]), (H, W, t)
        # the shadow is the bit string of the arrays: plane 0 = revealed
        bits = np.unpackbits(e[2].view(np.uint8), axis=1, bitorder="little")[:, :10 * HW]
        assert np.array_equal(bits.astype(bool), b["obs"].reshape(N, -1) != 0)


def test_delta_expansion_requires_both_arrays(oracle):
    from minesweeper_ppo_b200 import _lib
    L = _lib.load()
    desc = _lib.EnvDesc(16, 16, 40, 1, 0.0, 0.0, 0.0, 0, 0, 0)
    z = np.zeros((4, 8), np.int32)
    meta = np.zeros((4, 4), np.int32)
    sh = np.zeros((4, 41), np.uint64)
    obs = np.zeros((4, 10, 16, 16), np.float32)
    assert L.msw_expand_obs_host_delta(C.byref(desc), z.ctypes.data, z.ctypes.data, meta.ctypes.data, 4, obs.ctypes.data,
                                       None, sh.ctypes.data, 0, 1) != 0
    assert b"shadow" in L.msw_last_error()


@pytest.mark.filterwarnings("ignore:This process.*is multi-threaded:DeprecationWarning")
def test_expansion_works_in_a_forked_child(oracle):
    """The expander's worker threads do not survive fork(); a forked child (multiprocessing's default start method on
    Linux) must get a fresh pool instead of waiting for threads it does not have."""
    import os
    from minesweeper_ppo_b200 import _lib
    L = _lib.load()
    N, H, W, HW = 2000, 16, 16, 256
    rng = np.random.default_rng(1)
    mine = rng.random((N, HW)) < 0.15
    rev = (rng.random((N, HW)) < 0.4) & ~mine
    pm, pr = _pack(mine, HW), _pack(rev, HW)
    desc = _lib.EnvDesc(H, W, 40, 1, 0.0, 0.0, 0.0, 0, 0, 0)

    def expand():
        obs = np.empty((N, 10, H, W), np.float32)
        mask = np.empty((N, HW), bool)
        assert L.msw_expand_obs_host(C.byref(desc), pm.ctypes.data, pr.ctypes.data, None, N, obs.ctypes.data, mask.ctypes.data, 4) == 0
        return obs, mask

    want_obs, want_mask = expand()                       # the parent's pool now has worker threads
    pid = os.fork()
    if pid == 0:
        ok = 1
        try:
            import signal
            signal.alarm(20)                             # a child stuck on missing workers dies instead of hanging the suite
            o, m = expand()
            ok = 0 if (np.array_equal(o, want_obs) and np.array_equal(m, want_mask)) else 2
        finally:
            os._exit(ok)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0, status
    assert np.array_equal(want_mask, ~rev)


def test_worker_pool_under_changing_thread_counts_and_concurrent_callers(oracle):
    """The expander's worker pool (spin, then sleep; one parallel region at a time): thread counts that change from
    call to call (more threads than cores included), idle gaps long enough for the workers to fall asleep, and three
    Python threads calling at once -- every call still produces the full, correct result."""
    import threading
    import time
    from minesweeper_ppo_b200 import _lib
    L = _lib.load()
    N, H, W, HW = 1500, 16, 16, 256
    rng = np.random.default_rng(2)
    mine = rng.random((N, HW)) < 0.15
    rev = (rng.random((N, HW)) < 0.4) & ~mine
    pm, pr = _pack(mine, HW), _pack(rev, HW)
    desc = _lib.EnvDesc(H, W, 40, 1, 0.0, 0.0, 0.0, 0, 0, 0)

    def run(threads, obs, mask):
        assert L.msw_expand_obs_host(C.byref(desc), pm.ctypes.data, pr.ctypes.data, None, N, obs.ctypes.data,
                                     mask.ctypes.data, threads) == 0

    want, wmask = np.empty((N, 10, H, W), np.float32), np.empty((N, HW), bool)
    run(1, want, wmask)
    assert np.array_equal(wmask, ~rev)
    obs, mask = np.empty_like(want), np.empty_like(wmask)
    for k in range(300):
        obs.fill(-1.0)
        run((1, 2, 3, 8, 16, 5, 32)[k % 7], obs, mask)
        assert np.array_equal(obs, want) and np.array_equal(mask, wmask), k
        if k % 60 == 0:
            time.sleep(0.002)                                   # longer than the workers' spin: the wake-up path
    bad = []

    def caller():
        o, m = np.empty_like(want), np.empty_like(wmask)
        for k in range(100):
            run((4, 7, 2)[k % 3], o, m)
            if not (np.array_equal(o, want) and np.array_equal(m, wmask)):
                bad.append(k)

    threads = [threading.Thread(target=caller) for _ in range(3)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not bad
