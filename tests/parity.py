"""Shared parity harness: replay a golden fixture (tests/golden/*.npz, recorded from the
live reference by tests/golden/make_golden.py) through an env implementation and compare
every output bit for bit.  Used with the CPU oracle (-m "not gpu") and with the CUDA path
through the C-ABI (-m gpu)."""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ENV_CASES = [
    "16x16x40_valid", "16x16x40_any", "16x16x40_win", "16x30x99_valid", "16x30x99_win",
    "8x8x10_default", "4x4x8_fallback", "5x7x6_nosafe", "32x32x150_valid", "30x16x99_win",
]


def load(name: str):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unpack(bits: np.ndarray, n: int) -> np.ndarray:
    """Inverse of make_golden.pack on the last axis -> bool [..., n]."""
    return np.unpackbits(bits, axis=-1, count=n, bitorder="little").astype(bool)


def cfg_of(g) -> SimpleNamespace:
    return SimpleNamespace(
        H=int(g["H"]), W=int(g["W"]), mine_count=int(g["mine_count"]),
        guarantee_safe_neighborhood=bool(int(g["safe"])),
        win_reward=float(g["win_reward"]), loss_reward=float(g["loss_reward"]),
        step_penalty=float(g["step_penalty"]),
    )


def injections(g, t: int, N: int, HW: int):
    """Layouts the reference drew during step t -> (mine [N,HW] bool, sel [N] bool)."""
    sel = np.zeros(N, bool)
    mine = np.zeros((N, HW), bool)
    idx = np.nonzero(g["place_t"] == t)[0]
    if idx.size:
        env_i = g["place_i"][idx]
        sel[env_i] = True
        mine[env_i] = unpack(g["place_bits"][idx], HW)
    return mine, sel


def assert_bits_equal(a: np.ndarray, b: np.ndarray, what: str):
    """Bit-exact comparison (floats compared through their bit patterns)."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        assert b.dtype == a.dtype, f"{what}: dtype {a.dtype} vs {b.dtype}"
        a = a.view(f"u{a.dtype.itemsize}"); b = b.view(f"u{b.dtype.itemsize}")
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        raise AssertionError(f"{what}: {bad.shape[0]} mismatches, first at {bad[0].tolist()}")


def replay_env_case(name: str, make_env, check_state: bool = True):
    """make_env(cfg, N) -> adapter with reset()/step()/state() (see adapters below)."""
    g = load("env_" + name)
    cfg = cfg_of(g)
    N, T, HW = int(g["N"]), int(g["T"]), cfg.H * cfg.W
    env = make_env(cfg, N)
    obs, mask = env.reset()
    assert obs.dtype == np.float32 and obs.shape == (N, 10, cfg.H, cfg.W)
    assert_bits_equal(obs.reshape(N, -1), unpack(g["obs0"], 10 * HW).astype(np.float32), "reset obs")
    assert_bits_equal(mask, unpack(g["mask0"], HW), "reset mask")
    for t in range(T):
        mine, sel = injections(g, t, N, HW)
        o = env.step(g["actions"][t], mine, sel)
        tag = f"{name} t={t}"
        assert_bits_equal(o["obs"].reshape(N, -1), unpack(g["obs"][t], 10 * HW).astype(np.float32), tag + " obs")
        assert_bits_equal(o["mask"], unpack(g["mask"][t], HW), tag + " mask")
        assert_bits_equal(o["rewards"], g["rewards"][t], tag + " rewards")
        assert_bits_equal(o["dones"], g["dones"][t], tag + " dones")
        assert_bits_equal(o["outcome"], g["outcome"][t], tag + " outcome")
        assert_bits_equal(o["new_reveals"], g["new_reveals"][t], tag + " last_new_reveals")
        assert_bits_equal(o["step"], g["step"][t], tag + " aux.step")
        frac = o["revealed_count"].astype(np.int64) / max(1, HW)          # env.py:165
        assert_bits_equal(frac.astype(np.float64), g["revealed_frac"][t], tag + " revealed_frac")
        assert_bits_equal(o["labels"].reshape(N, -1), unpack(g["mine_labels"][t], HW).astype(np.float32), tag + " mine_labels")
        assert_bits_equal(o["valid"].reshape(N, -1), unpack(g["mine_valid"][t], HW), tag + " mine_valid")
        if check_state:
            s = env.state()
            assert_bits_equal(s["revealed"], unpack(g["st_revealed"][t], HW), tag + " state.revealed")
            assert_bits_equal(s["mine"], unpack(g["st_mine"][t], HW), tag + " state.mine_mask")
            assert_bits_equal(s["counts"], g["st_counts"][t], tag + " state.adjacent_counts")
            assert_bits_equal(s["first"], g["st_first"][t], tag + " state.first_click_done")
            assert_bits_equal(s["step_count"], g["st_step_count"][t], tag + " state.step_count")
    return g


class OracleAdapter:
    def __init__(self, O, cfg, N, nthreads=1):
        self.v = O.OracleVecEnv(N, cfg, seed=0, nthreads=nthreads, aux_maps=True)

    def reset(self):
        b = self.v.reset()
        return b["obs"], b["action_mask"]

    def step(self, actions, mine, sel):
        b, r, d, info = self.v.step(actions, mine, sel, tensor_infos=True)
        return dict(obs=b["obs"], mask=b["action_mask"], rewards=r, dones=d,
                    outcome=info["outcome_code"], new_reveals=info["last_new_reveals"],
                    step=info["step"], revealed_count=info["revealed_count"],
                    labels=self.v.mine_labels, valid=self.v.mine_valid)

    def state(self):
        v = self.v
        return dict(revealed=v.revealed.astype(bool), mine=v.mine.astype(bool), counts=v.counts.copy(),
                    first=v.first_click_done.astype(bool), step_count=v.step_count.copy())


class CudaAdapter:
    """The product path: minesweeper_ppo_b200.VecMinesweeper through the C ABI (native torch API)."""

    def __init__(self, cfg, N, seed=0, env_id_base=0):
        import torch
        import minesweeper_ppo_b200 as m
        self.torch = torch
        ec = m.EnvConfig(H=cfg.H, W=cfg.W, mine_count=cfg.mine_count,
                         guarantee_safe_neighborhood=cfg.guarantee_safe_neighborhood,
                         win_reward=cfg.win_reward, loss_reward=cfg.loss_reward, step_penalty=cfg.step_penalty)
        self.v = m.VecMinesweeper(N, ec, seed=seed, api="torch", aux_maps=True, env_id_base=env_id_base)

    def reset(self):
        b = self.v.reset()
        return b["obs"].cpu().numpy(), b["action_mask"].cpu().numpy()

    def step(self, actions, mine=None, sel=None):
        if sel is not None:
            self.v.inject_layouts(mine, sel)
        a = actions if isinstance(actions, self.torch.Tensor) else self.torch.from_numpy(np.ascontiguousarray(actions))
        b, r, d, info = self.v.step(a)
        c = lambda t: t.cpu().numpy()
        return dict(obs=c(b["obs"]), mask=c(b["action_mask"]), rewards=c(r), dones=c(d),
                    outcome=c(info["outcome_code"]), new_reveals=c(info["last_new_reveals"]),
                    step=c(info["step"]), revealed_count=c(info["revealed_count"]),
                    labels=c(self.v.mine_labels), valid=c(self.v.mine_valid))

    def state(self):
        u = self.v._unpacked()
        return dict(revealed=u["revealed"].astype(bool), mine=u["mine"].astype(bool), counts=u["counts"].copy(),
                    first=u["meta"][:, 0].astype(bool), step_count=u["meta"][:, 1].copy())


def scripted_policy(H: int, W: int, seed: int = 0):
    """A deterministic stand-in for the policy network whose outputs are EXACT small integers /
    quarter-integers in fp32 on any device (shifted adds only, no convolution kernels), so a greedy
    evaluation gives the same trajectories on the reference (CPU) and on the CUDA env.  It prefers
    frontier cells with small neighbouring numbers, so games run long enough to exercise the analytics."""
    import torch

    class ScriptedPolicy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(seed)
            self.table = torch.nn.Parameter(torch.randperm(H * W, generator=g).float(), requires_grad=False)

        @staticmethod
        def _sum3x3(x):                      # [B,1,H,W] -> sum over the 3x3 neighbourhood, zero padded
            p = torch.nn.functional.pad(x, (1, 1, 1, 1))
            out = torch.zeros_like(x)
            for dr in range(3):
                for dc in range(3):
                    out = out + p[:, :, dr:dr + H, dc:dc + W]
            return out

        def forward(self, obs, return_mine: bool = False):
            B = obs.shape[0]
            rev = obs[:, 0:1]
            num = torch.zeros_like(rev)
            for k in range(1, 9):
                num = num + float(k) * obs[:, 1 + k:2 + k]
            nb, ns = self._sum3x3(rev), self._sum3x3(num)
            logits = (1024.0 * (nb > 0).float() - 16.0 * ns).reshape(B, H * W) + self.table
            value = torch.zeros((B,), dtype=obs.dtype, device=obs.device)
            if not return_mine:
                return logits, value
            return logits, value, (ns - 2.0 * nb) * 0.25

    return ScriptedPolicy()
