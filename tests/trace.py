"""At-scale traces of the reference env (BASELINE.md section 4, C2 parity sub-run: N=4,096 x 256 steps).

Three action series, all derived deterministically from a seeded NumPy generator and the env's own
outputs (so nothing but the digests needs storing):

  A  uniformly random VALID cell per env per step   (s=rng.random(mask.shape); s[~mask]=-1; argmax)
  B  uniformly random ANY cell in [-HW, 2HW)        (no-op clicks, Python modulo of env.py:106)
  C  "careful" play: with probability 0.97 a random SAFE hidden cell, else a random valid cell --
     long episodes, deep flood fills and thousands of WINS (random play never wins at 16x16x40)

The same driver loop runs against
  * the live reference (`ReferenceSide`, needs baseline/_ref): used by tests/golden/make_trace.py to
    record per-step SHA-256 digests of every output and of the per-env state, and by the -m gpu live
    test to compare arrays directly, and
  * an implementation under test (`parity.OracleAdapter` / `CudaAdapter`), which is fed the
    reference's mine layouts: captured live, or -- where the reference is absent -- REGENERATED from
    NumPy's own PCG64 streams by `LayoutReplayer`, a restatement of env.py:280-312 / :393-395 in test
    code (the digests come from the real reference, so a wrong restatement cannot pass).

Test infrastructure only.
"""
from __future__ import annotations

import hashlib
import os
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FIELDS = ("obs", "mask", "rewards", "dones", "outcome", "new_reveals", "step", "revealed_frac",
          "labels", "valid", "st_revealed", "st_mine", "st_counts", "st_first", "st_step_count")
DTYPES = dict(obs=np.float32, mask=bool, rewards=np.float32, dones=bool, outcome=np.int8, new_reveals=np.int32,
              step=np.int32, revealed_frac=np.float64, labels=np.float32, valid=bool, st_revealed=bool,
              st_mine=bool, st_counts=np.uint8, st_first=bool, st_step_count=np.int32)
CAREFUL_P = 0.97
OUTCOME = {None: 0, "win": 1, "loss": 2}


def trace_cfg(H=16, W=16, mines=40) -> SimpleNamespace:
    return SimpleNamespace(H=H, W=W, mine_count=mines, guarantee_safe_neighborhood=True,
                           win_reward=1.0, loss_reward=-1.0, step_penalty=1e-4)


def choose_actions(series: str, rng: np.random.Generator, mask: np.ndarray, cur_mine: np.ndarray,
                   fresh: np.ndarray) -> np.ndarray:
    N, HW = mask.shape
    if series == "A":
        s = rng.random(mask.shape)
        s[~mask] = -1.0
        return s.argmax(1).astype(np.int64)
    if series == "B":
        return rng.integers(-HW, 2 * HW, size=N, dtype=np.int64)
    if series == "C":
        s = rng.random(mask.shape)
        safe = mask & ~cur_mine
        careful = (rng.random(N) < CAREFUL_P) & ~fresh & safe.any(1)
        elig = np.where(careful[:, None], safe, mask)
        s[~elig] = -1.0
        return s.argmax(1).astype(np.int64)
    raise ValueError(series)


def digest(name: str, a: np.ndarray) -> bytes:
    a = np.ascontiguousarray(np.asarray(a), dtype=DTYPES[name])
    return hashlib.sha256(a.view(np.uint8).reshape(-1).data).digest()


def normalise(out: Dict[str, np.ndarray], N: int, HW: int) -> Dict[str, np.ndarray]:
    """Bring a step result (reference side or adapter side) to the canonical shapes/dtypes of FIELDS."""
    o = {}
    for k in FIELDS:
        a = np.asarray(out[k])
        if a.dtype != DTYPES[k]:
            b = a.astype(DTYPES[k])
            assert np.array_equal(b.astype(a.dtype), a), f"{k}: lossy cast {a.dtype}->{DTYPES[k]}"
            a = b
        o[k] = a.reshape(N, -1) if a.ndim > 1 else a
    return o


class LayoutReplayer:
    """The mine layouts the reference WOULD draw: VecMinesweeper seeds one PCG64 generator per env from
    `default_rng(seed).integers(0, 2**31-1, size=N, dtype=int64)` (env.py:393-395, :49) and every first
    click consumes `rng.choice(allowed_indices, size=mines, replace=False)` (env.py:309), allowed = all
    cells but the clicked one and (if guarantee_safe_neighborhood) its clipped 3x3, relaxed to the
    clicked cell alone when fewer than `mines` cells remain (env.py:286-307)."""

    def __init__(self, N: int, cfg, seed: int):
        base = np.random.default_rng(seed)
        seeds = base.integers(0, 2**31 - 1, size=N, dtype=np.int64)
        self.gens = [np.random.default_rng(int(s)) for s in seeds]
        self.cfg, self.H, self.W = cfg, int(cfg.H), int(cfg.W)

    def place(self, i: int, cell: int) -> np.ndarray:
        H, W, M = self.H, self.W, int(self.cfg.mine_count)
        r0, c0 = divmod(int(cell), W)
        forbidden = np.zeros((H, W), bool)
        if self.cfg.guarantee_safe_neighborhood:
            forbidden[max(0, r0 - 1):r0 + 2, max(0, c0 - 1):c0 + 2] = True
        forbidden[r0, c0] = True
        allowed = np.flatnonzero(~forbidden)
        if len(allowed) < M:
            forbidden[:] = False
            forbidden[r0, c0] = True
            allowed = np.flatnonzero(~forbidden)
        pos = self.gens[i].choice(allowed, size=M, replace=False)
        mine = np.zeros(H * W, bool)
        mine[pos] = True
        return mine


class ReferenceSide:
    """The live reference: VecMinesweeper(N, cfg, seed) with _place_mines_safe wrapped (not edited)."""

    def __init__(self, N: int, cfg, seed: int):
        import reference_live as RL
        mods = RL.load()
        E = mods["env"]
        self.rec = RL.LayoutRecorder()
        self.rec.__enter__()
        self.cfg = E.EnvConfig(H=cfg.H, W=cfg.W, mine_count=cfg.mine_count,
                               guarantee_safe_neighborhood=cfg.guarantee_safe_neighborhood,
                               win_reward=cfg.win_reward, loss_reward=cfg.loss_reward, step_penalty=cfg.step_penalty)
        self.vec = E.VecMinesweeper(N, self.cfg, seed=seed)
        self.index_of = {id(e): i for i, e in enumerate(self.vec.envs)}
        self.N, self.HW = N, cfg.H * cfg.W

    def close(self):
        self.rec.__exit__(None, None, None)

    def reset(self) -> Tuple[np.ndarray, np.ndarray]:
        b = self.vec.reset()
        assert not self.rec.log
        return b["obs"], b["action_mask"]

    def step(self, actions: np.ndarray):
        """-> (canonical outputs, mine [N,HW] bool, sel [N] bool) with the layouts placed in this step."""
        N, HW, H, W = self.N, self.HW, self.cfg.H, self.cfg.W
        batch, rew, done, infos = self.vec.step(actions)
        mine, sel = np.zeros((N, HW), bool), np.zeros(N, bool)
        for i, m in self.rec.drain(self.index_of):
            assert not sel[i]
            sel[i], mine[i] = True, m.reshape(-1)
        envs = self.vec.envs
        first = np.array([e.first_click_done for e in envs], bool)
        st_mine = np.stack([e.mine_mask for e in envs]).reshape(N, HW)
        st_rev = np.stack([e.revealed for e in envs]).reshape(N, HW)
        st_flags = np.stack([e.flags for e in envs]).reshape(N, HW)
        # auxiliary maps exactly as train_rl.py:205-212 derives them from vec.envs
        labels = np.where(first[:, None], st_mine, False).astype(np.float32)
        valid = np.where(first[:, None], ~st_rev & ~st_flags, False)
        out = dict(
            obs=batch["obs"], mask=batch["action_mask"], rewards=rew, dones=done,
            outcome=np.array([OUTCOME[o] for o in infos["outcome"]], np.int8),
            new_reveals=np.array([x["last_new_reveals"] for x in infos["aux"]], np.int32),
            step=np.array([x["step"] for x in infos["aux"]], np.int32),
            revealed_frac=np.array([x["revealed_frac"] for x in infos["aux"]], np.float64),
            labels=labels, valid=valid, st_revealed=st_rev, st_mine=st_mine,
            st_counts=np.stack([e.adjacent_counts for e in envs]).reshape(N, HW),
            st_first=first, st_step_count=np.array([e.step_count for e in envs], np.int32),
        )
        assert infos["done"] == [bool(d) for d in done]
        assert batch["obs"].dtype == np.float32 and rew.dtype == np.float32
        return normalise(out, N, HW), mine, sel


def adapter_step(env, actions, mine, sel, N: int, HW: int) -> Dict[str, np.ndarray]:
    """One step of a tests.parity adapter (OracleAdapter / CudaAdapter) in canonical form."""
    o = env.step(actions, mine, sel)
    s = env.state()
    frac = o["revealed_count"].astype(np.int64) / max(1, HW)          # env.py:165
    out = dict(obs=o["obs"], mask=o["mask"], rewards=o["rewards"], dones=o["dones"], outcome=o["outcome"],
               new_reveals=o["new_reveals"], step=o["step"], revealed_frac=frac.astype(np.float64),
               labels=o["labels"], valid=o["valid"], st_revealed=s["revealed"], st_mine=s["mine"],
               st_counts=s["counts"], st_first=s["first"], st_step_count=s["step_count"])
    assert o["obs"].dtype == np.float32 and o["rewards"].dtype == np.float32
    return normalise(out, N, HW)


class Driver:
    """Action selection shared by both sides: keeps the harness's knowledge of the boards (which env is
    fresh, which layout each env currently has) from the layouts it injects / observes."""

    def __init__(self, series: str, N: int, HW: int, seed: int):
        self.series, self.N, self.HW = series, N, HW
        self.rng = np.random.default_rng(seed)
        self.cur_mine = np.zeros((N, HW), bool)
        self.fresh = np.ones(N, bool)

    def actions(self, mask: np.ndarray) -> np.ndarray:
        return choose_actions(self.series, self.rng, mask, self.cur_mine, self.fresh)

    def observe(self, mine: np.ndarray, sel: np.ndarray, dones: np.ndarray):
        self.cur_mine[sel] = mine[sel]
        self.fresh = np.asarray(dones, bool).copy()


def fixture_path(series: str, N: int, T: int, cfg) -> str:
    return os.path.join(GOLDEN, f"trace_{cfg.H}x{cfg.W}x{cfg.mine_count}_{series}_{N}x{T}.npz")


def record_reference_trace(series: str, N: int, T: int, cfg, env_seed: int = 0, action_seed: int = 1) -> Dict:
    """Run the live reference and return the fixture dict (digests [T, len(FIELDS), 32] u8 + summary)."""
    HW = cfg.H * cfg.W
    ref = ReferenceSide(N, cfg, env_seed)
    try:
        obs0, mask = ref.reset()
        drv = Driver(series, N, HW, action_seed)
        dig = np.zeros((T, len(FIELDS), 32), np.uint8)
        wins = losses = placed = noop = 0
        max_new = 0
        for t in range(T):
            a = drv.actions(mask)
            out, mine, sel = ref.step(a.astype(np.int32) if series != "B" else a)
            for f, k in enumerate(FIELDS):
                dig[t, f] = np.frombuffer(digest(k, out[k]), np.uint8)
            wins += int((out["outcome"] == 1).sum()); losses += int((out["outcome"] == 2).sum())
            placed += int(sel.sum()); max_new = max(max_new, int(out["new_reveals"].max()))
            noop += int(((out["new_reveals"] == 0) & (out["outcome"] == 0)).sum())
            drv.observe(mine, sel, out["dones"])
            mask = out["mask"].astype(bool)
        return dict(digests=dig, fields=np.array(FIELDS), N=N, T=T, H=cfg.H, W=cfg.W, mine_count=cfg.mine_count,
                    env_seed=env_seed, action_seed=action_seed, series=series, wins=wins, losses=losses,
                    layouts_placed=placed, noop_clicks=noop, max_new_reveals=max_new,
                    obs0_digest=np.frombuffer(digest("obs", obs0), np.uint8))
    finally:
        ref.close()


def replay_trace(series: str, N: int, T: int, cfg, make_env, steps: Optional[int] = None) -> Dict[str, int]:
    """Replay a recorded trace through `make_env(cfg, N)` (a tests.parity adapter), layouts regenerated
    by LayoutReplayer, and compare every per-step digest.  `steps` < T replays a prefix."""
    g = np.load(fixture_path(series, N, T, cfg))
    assert tuple(g["fields"]) == FIELDS
    HW = cfg.H * cfg.W
    env = make_env(cfg, N)
    obs0, mask = env.reset()
    assert digest("obs", obs0) == g["obs0_digest"].tobytes(), "reset obs"
    drv = Driver(series, N, HW, int(g["action_seed"]))
    lay = LayoutReplayer(N, cfg, int(g["env_seed"]))
    wins = placed = 0
    for t in range(T if steps is None else min(T, steps)):
        a = drv.actions(mask)
        sel = drv.fresh.copy()
        mine = np.zeros((N, HW), bool)
        cells = np.mod(a, HW)
        for i in np.nonzero(sel)[0]:
            mine[i] = lay.place(int(i), int(cells[i]))
        out = adapter_step(env, a.astype(np.int32) if series != "B" else a, mine, sel, N, HW)
        for f, k in enumerate(FIELDS):
            if digest(k, out[k]) != g["digests"][t, f].tobytes():
                raise AssertionError(f"series {series} step {t}: {k} differs from the reference trace "
                                     f"({fixture_path(series, N, T, cfg)})")
        wins += int((out["outcome"] == 1).sum()); placed += int(sel.sum())
        drv.observe(mine, sel, out["dones"])
        mask = out["mask"].astype(bool)
    return dict(wins=wins, layouts_placed=placed)


def lockstep_live(series: str, N: int, T: int, cfg, make_env, env_seed: int = 0, action_seed: int = 1) -> Dict[str, int]:
    """The live reference and an implementation side by side: every output and the per-env state of
    every step compared array against array (bit patterns for floats)."""
    from parity import assert_bits_equal
    HW = cfg.H * cfg.W
    ref = ReferenceSide(N, cfg, env_seed)
    try:
        env = make_env(cfg, N)
        obs0, mask = ref.reset()
        o2, m2 = env.reset()
        assert_bits_equal(o2, obs0, "reset obs"); assert_bits_equal(m2, mask, "reset mask")
        drv = Driver(series, N, HW, action_seed)
        wins = losses = placed = 0
        for t in range(T):
            a = drv.actions(mask)
            a = a.astype(np.int32) if series != "B" else a
            want, mine, sel = ref.step(a)
            assert np.array_equal(sel, drv.fresh), f"t={t}: the reference placed mines on a non-fresh board?"
            got = adapter_step(env, a, mine, sel, N, HW)
            for k in FIELDS:
                assert_bits_equal(got[k], want[k], f"series {series} t={t} {k}")
            wins += int((want["outcome"] == 1).sum()); losses += int((want["outcome"] == 2).sum())
            placed += int(sel.sum())
            drv.observe(mine, sel, want["dones"])
            mask = want["mask"].astype(bool)
        return dict(wins=wins, losses=losses, layouts_placed=placed)
    finally:
        ref.close()
