"""avoidability.analyze_avoidability (avoidability.py:145-394): oracle vs
fixtures recorded from the live reference (CPU); CUDA kernel vs the same fixtures and vs the oracle
(GPU).  Sets, lists and flags are compared exactly."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import parity as P


def _cases():
    g = P.load("avoidability")
    for name in g["names"]:
        name = str(name)
        H, W, M = (int(x) for x in g[f"{name}_cfg"])
        HW = H * W
        yield NS(name=name, H=H, W=W, M=M, HW=HW, rev=P.unpack(g[f"{name}_rev"], HW), mine=P.unpack(g[f"{name}_mine"], HW),
                 fcd=g[f"{name}_fcd"].astype(bool), chosen=g[f"{name}_chosen"], avoidable=g[f"{name}_avoidable"].astype(bool),
                 safe=P.unpack(g[f"{name}_safe"], HW), ncomp=g[f"{name}_ncomp"], sizes=g[f"{name}_sizes"],
                 cfs=g[f"{name}_cfs"].astype(bool), ccs=g[f"{name}_ccs"])


def result_fields(arr, i, chosen, revealed_i):
    """(avoidable, forced_safe set, component_sizes, chosen_is_forced_safe, chosen_component_size) of env i
    from the array form both implementations emit (oracle/msw_oracle.h, include/msw_b200.h)."""
    fl = int(arr["flags"][i])
    sizes = [int(s) for s in arr["comp_size"][i] if s > 0]
    safe = {int(k) for k in np.flatnonzero(arr["safe"][i])}
    cfs, ccs = False, None
    if chosen is not None and chosen >= 0:
        if fl & 2:                                        # frontier exists (avoidability.py:239-249)
            lab = int(arr["comp_of_cell"][i][chosen])
            if lab >= 0:
                ccs = int(arr["comp_size"][i][lab])
                cfs = chosen in safe
        elif fl & 4:                                      # first click done, empty frontier (:176-186)
            ccs = None if revealed_i[chosen] else 1
    return bool(fl & 1), safe, sizes, cfs, ccs


def check_against_fixture(c, arr, what):
    for i in range(c.rev.shape[0]):
        av, safe, sizes, cfs, ccs = result_fields(arr, i, int(c.chosen[i]), c.rev[i])
        tag = f"{what} {c.name} state {i}"
        assert av == bool(c.avoidable[i]), tag
        assert safe == {int(k) for k in np.flatnonzero(c.safe[i])}, tag
        assert sizes == [int(s) for s in c.sizes[i][: int(c.ncomp[i])]], tag
        assert cfs == bool(c.cfs[i]), tag
        assert (-1 if ccs is None else ccs) == int(c.ccs[i]), tag


def _oracle_env(oracle, c):
    n = c.rev.shape[0]
    cfg = NS(H=c.H, W=c.W, mine_count=c.M, guarantee_safe_neighborhood=True, win_reward=1.0, loss_reward=-1.0,
             step_penalty=1e-4)
    v = oracle.OracleVecEnv(n, cfg)
    v.revealed[:] = c.rev; v.mine[:] = c.mine; v.first_click_done[:] = c.fcd
    for i in range(n):
        v.counts[i] = oracle.adjacent_counts(c.mine[i].reshape(c.H, c.W)).reshape(-1)
    return v


def test_oracle_avoidability_matches_reference(oracle):
    states = safe_cells = 0
    for c in _cases():
        check_against_fixture(c, _oracle_env(oracle, c).avoidability(), "oracle")
        states += c.rev.shape[0]; safe_cells += int(c.safe.sum())
    assert states >= 2000 and safe_cells > 10000


def test_oracle_avoidability_respects_flags(oracle):
    """Flagged hidden cells are neither frontier variables nor masked out of the counts
    (avoidability.py:162, 208): a flag on the only unknown neighbour removes the constraint's variable."""
    H = W = 4
    mine = np.zeros((1, 16), bool); mine[0, 5] = True
    rev = np.zeros((1, 16), bool); rev[0, [0, 1, 2, 4, 8]] = True
    cfg = NS(H=H, W=W, mine_count=1, guarantee_safe_neighborhood=True, win_reward=1.0, loss_reward=-1.0, step_penalty=1e-4)
    v = oracle.OracleVecEnv(1, cfg)
    v.revealed[:] = rev; v.mine[:] = mine; v.first_click_done[:] = 1
    v.counts[0] = oracle.adjacent_counts(mine[0].reshape(H, W)).reshape(-1)
    plain = v.avoidability()
    assert plain["comp_of_cell"][0][5] >= 0
    v.flags[0, 5] = 1
    flagged = v.avoidability()
    assert flagged["comp_of_cell"][0][5] == -1 and not flagged["safe"][0][5]


@pytest.mark.gpu
def test_cuda_avoidability_matches_reference_and_oracle(oracle):
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200.avoidability import analyze_avoidability
    for c in _cases():
        n = c.rev.shape[0]
        v = m.VecMinesweeper(n, m.EnvConfig(H=c.H, W=c.W, mine_count=c.M), api="torch")
        v.reset()
        v.set_state(mine=c.mine, revealed=c.rev, first_click_done=c.fcd.astype(np.int32))
        arr = v.avoidability()
        assert not (arr["flags"] & 8).any(), "search budget exceeded"
        check_against_fixture(c, arr, "cuda")
        want = _oracle_env(oracle, c).avoidability()
        for k in ("safe", "comp_of_cell", "comp_size"):
            P.assert_bits_equal(arr[k], want[k], f"cuda vs oracle {c.name} {k}")
        P.assert_bits_equal(arr["flags"] & 7, want["flags"], f"cuda vs oracle {c.name} flags")
        for i in (0, n // 3, n - 1):                      # reference call shape on a vec.envs[i] view
            ch = int(c.chosen[i])
            res = analyze_avoidability(v.envs[i], None if ch < 0 else ch)
            assert res.avoidable == bool(c.avoidable[i])
            assert res.forced_safe_cells == {int(k) for k in np.flatnonzero(c.safe[i])}
            assert res.component_sizes == [int(s) for s in c.sizes[i][: int(c.ncomp[i])]]
            assert res.chosen_is_forced_safe == bool(c.cfs[i])
            assert (-1 if res.chosen_component_size is None else res.chosen_component_size) == int(c.ccs[i])
            assert res.count_forced_safe_cells == int(c.safe[i].sum())


@pytest.mark.gpu
def test_cuda_avoidability_with_flags_matches_oracle(oracle):
    """Random flags on hidden SAFE cells (the reference env never sets flags itself, but the analysis
    reads them, avoidability.py:162): CUDA == oracle on flagged states.  Flags on mines would make the
    constraint system inconsistent (a flagged cell is no variable but still counts in the numbers), and
    the reference's deductions on an inconsistent system depend on its iteration order."""
    import minesweeper_ppo_b200 as m
    rng = np.random.default_rng(11)
    for c in _cases():
        if c.name not in ("8x8x10", "9x9x30_dense", "16x16x40"):
            continue
        n = c.rev.shape[0]
        flags = (~c.rev) & (~c.mine) & (rng.random(c.rev.shape) < 0.08)
        o = _oracle_env(oracle, c)
        o.flags[:] = flags
        v = m.VecMinesweeper(n, m.EnvConfig(H=c.H, W=c.W, mine_count=c.M), api="torch")
        v.reset()
        v.set_state(mine=c.mine, revealed=c.rev, flags=flags, first_click_done=c.fcd.astype(np.int32))
        got, want = v.avoidability(), o.avoidability()
        for k in ("safe", "comp_of_cell", "comp_size"):
            P.assert_bits_equal(got[k], want[k], f"flags {c.name} {k}")
        P.assert_bits_equal(got["flags"] & 7, want["flags"], f"flags {c.name} flags")
