"""rules.analyze_forced_modules (rules.py:206-259): oracle vs fixtures recorded
from the reference (CPU), CUDA kernel vs the same fixtures (GPU).  Sets of cells, compared exactly."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import parity as P


def _cases():
    g = P.load("forced_subset")
    for name in g["names"]:
        H, W, M = (int(x) for x in g[f"{name}_cfg"])
        HW = H * W
        yield (str(name), H, W, M, P.unpack(g[f"{name}_rev"], HW), P.unpack(g[f"{name}_mine"], HW),
               P.unpack(g[f"{name}_subset"], HW))


def test_oracle_forced_subset_matches_reference(oracle):
    total = 0
    for name, H, W, M, rev, mine, want in _cases():
        n = rev.shape[0]
        cfg = NS(H=H, W=W, mine_count=M, guarantee_safe_neighborhood=True, win_reward=1.0, loss_reward=-1.0,
                 step_penalty=1e-4)
        v = oracle.OracleVecEnv(n, cfg)
        v.revealed[:] = rev; v.mine[:] = mine
        for i in range(n):
            v.counts[i] = oracle.adjacent_counts(mine[i].reshape(H, W)).reshape(-1)
        got = v.forced_subset()
        P.assert_bits_equal(got, want, f"forced_subset {name}")
        total += int(want.sum())
    assert total > 5000           # the fixtures are not trivially empty


@pytest.mark.gpu
def test_cuda_forced_subset_matches_reference():
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200.rules import analyze_forced_modules
    for name, H, W, M, rev, mine, want in _cases():
        n = rev.shape[0]
        v = m.VecMinesweeper(n, m.EnvConfig(H=H, W=W, mine_count=M), api="torch")
        v.reset()
        v.set_state(mine=mine, revealed=rev, first_click_done=np.ones(n, np.int32))
        P.assert_bits_equal(v.forced_subset(), want, f"forced_subset {name}")
        for i in (0, n // 2, n - 1):           # reference call shape on a vec.envs[i] view
            assert analyze_forced_modules(v.envs[i]) == {"subset_reveal": {int(k) for k in np.flatnonzero(want[i])}}
