"""-m gpu: property test of the flood-fill reveal: for arbitrary boards -- random shape, mine
density, already-revealed set, FLAGS (which the hot path never sets but the reference's flood fill honours,
env_numba.py:50-75) and start cell -- one CUDA step equals the oracle's array-queue BFS, and the result is a
fixed point: every newly revealed zero-count cell has all its unflagged, non-mine neighbours revealed."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

SHAPES = [(16, 16), (16, 30), (30, 16), (9, 9), (5, 7), (32, 32), (1, 32), (24, 1), (3, 31)]


@settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
@given(shape=st.sampled_from(SHAPES), mine_p=st.floats(0.0, 0.35), rev_p=st.floats(0.0, 0.6), flag_p=st.floats(0.0, 0.3),
       seed=st.integers(0, 2**31 - 1))
def test_flood_fill_equals_oracle_on_random_boards(oracle, shape, mine_p, rev_p, flag_p, seed):
    import torch
    import minesweeper_ppo_b200 as m
    H, W = shape
    HW, n = H * W, 96
    rng = np.random.default_rng(seed)
    mines = rng.random((n, HW)) < mine_p
    rev = (rng.random((n, HW)) < rev_p) & ~mines                 # revealed cells are never mines in a live game
    flags = (rng.random((n, HW)) < flag_p) & ~rev
    cells = rng.integers(0, HW, size=n)
    v = m.VecMinesweeper(n, m.EnvConfig(H=H, W=W, mine_count=0), api="torch")
    v.reset()
    v.set_state(mine=mines, revealed=rev, flags=flags, first_click_done=np.ones(n, np.int32))
    _, _, dones, info = v.step(torch.from_numpy(cells.astype(np.int64)).cuda())
    new = info["last_new_reveals"].cpu().numpy()
    dones = dones.cpu().numpy()
    after = v._unpacked()["revealed"].astype(bool)
    for k in range(n):
        r, c = divmod(int(cells[k]), W)
        counts = oracle.adjacent_counts(mines[k].reshape(H, W))
        if rev[k, cells[k]]:
            assert new[k] == 0 and not dones[k]                   # already open: no-op (env.py:138-140)
            continue
        if mines[k, cells[k]]:
            assert dones[k]                                       # loss: the flood fill is never entered
            continue
        want, cnt = oracle.flood_fill(rev[k].reshape(H, W), flags[k].reshape(H, W), mines[k].reshape(H, W), counts, r, c)
        assert new[k] == cnt, (shape, k)
        if dones[k]:
            continue                                              # a win auto-resets the board
        got = after[k].reshape(H, W)
        assert np.array_equal(got, want), (shape, k)
        # fixed point: newly revealed zero cells have every eligible neighbour revealed
        newly = got & ~rev[k].reshape(H, W)
        zr, zc = np.nonzero(newly & (counts == 0))
        for rr, cc in zip(zr, zc):
            for dr in (-1, 0, 1):
                for dc in (-1, 0, 1):
                    y, x = rr + dr, cc + dc
                    if 0 <= y < H and 0 <= x < W and not flags[k].reshape(H, W)[y, x] and not mines[k].reshape(H, W)[y, x]:
                        assert got[y, x]
