"""ctypes front-end of the CPU oracle (oracle/msw_oracle.c).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product package
(minesweeper_ppo_b200), which has no CPU fallback.

`OracleVecEnv` mirrors the reference's `VecMinesweeper` (minesweeper/env.py:379-517)
closely enough that parity tests read like reference usage: NumPy in, NumPy out,
`reset() -> {"obs","action_mask"}`, `step(actions) -> (batch, rewards, dones, infos)`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Any, Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsw_oracle.so")
_lib = None

OBS_CHANNELS = 10  # env.py:79-85


def build(force: bool = False) -> str:
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ("msw_oracle.c", "msw_oracle_avoid.c", "msw_oracle.h", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmsw_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Cfg(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("mine_count", C.c_int32), ("safe_nbhd", C.c_int32),
        ("win_reward", C.c_double), ("loss_reward", C.c_double), ("step_penalty", C.c_double),
        ("seed", C.c_uint64),
    ]


class _State(C.Structure):
    _fields_ = [
        ("mine", C.c_void_p), ("revealed", C.c_void_p), ("flags", C.c_void_p), ("counts", C.c_void_p),
        ("first_click_done", C.c_void_p), ("step_count", C.c_void_p),
        ("last_new_reveals", C.c_void_p), ("episode_idx", C.c_void_p),
    ]


class _StepOut(C.Structure):
    _fields_ = [
        ("obs", C.c_void_p), ("mask", C.c_void_p), ("reward", C.c_void_p), ("done", C.c_void_p),
        ("outcome", C.c_void_p), ("new_reveals", C.c_void_p), ("step", C.c_void_p),
        ("revealed_count", C.c_void_p), ("mine_labels", C.c_void_p), ("mine_valid", C.c_void_p),
    ]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.orc_version.restype = C.c_int
        L.orc_philox4x32_10.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_place_mines.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.orc_adjacent_counts.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_flood_fill.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int]
        L.orc_flood_fill.restype = C.c_int
        L.orc_encode.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4
        L.orc_vec_reset.argtypes = [C.POINTER(_Cfg), C.c_int64, C.POINTER(_State)] + [C.c_void_p] * 4 + [C.c_int]
        L.orc_vec_step.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.POINTER(_State), C.POINTER(_StepOut), C.c_int]
        L.orc_vec_encode.argtypes = [C.POINTER(_Cfg), C.c_int64, C.POINTER(_State)] + [C.c_void_p] * 4 + [C.c_int]
        L.orc_late_start.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_int64, C.POINTER(_State), C.c_void_p, C.c_uint64,
                                     C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_forced_subset.argtypes = [C.POINTER(_Cfg), C.c_int64, C.POINTER(_State), C.c_void_p]
        L.orc_avoidability.argtypes = [C.POINTER(_Cfg), C.c_int64, C.POINTER(_State)] + [C.c_void_p] * 4
        L.orc_gae.argtypes = [C.c_int64, C.c_int64] + [C.c_void_p] * 4 + [C.c_float, C.c_float] + [C.c_void_p] * 2
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


@dataclass
class OracleEnvConfig:
    """Same fields and defaults as the reference EnvConfig (env.py:19-30)."""
    H: int = 8
    W: int = 8
    mine_count: int = 10
    guarantee_safe_neighborhood: bool = True
    use_pair_constraints: Optional[bool] = None
    solver_preset: str = "zf"
    win_reward: float = 1.0
    loss_reward: float = -1.0
    step_penalty: float = 1e-4


def _ccfg(cfg, seed: int) -> _Cfg:
    return _Cfg(int(cfg.H), int(cfg.W), int(cfg.mine_count), int(bool(cfg.guarantee_safe_neighborhood)),
                float(cfg.win_reward), float(cfg.loss_reward), float(cfg.step_penalty),
                int(seed) & 0xFFFFFFFFFFFFFFFF)


def philox4x32_10(key: Tuple[int, int], ctr: Tuple[int, int, int, int]) -> Tuple[int, ...]:
    c = np.array(ctr, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(key[0], key[1], c.ctypes.data, o.ctypes.data)
    return tuple(int(x) for x in o)


def place_mines(cfg, seed: int, env_id: int, episode: int, r0: int, c0: int) -> np.ndarray:
    out = np.zeros((cfg.H, cfg.W), dtype=np.uint8)
    cc = _ccfg(cfg, seed)
    lib().orc_place_mines(C.byref(cc), env_id, episode, r0, c0, out.ctypes.data)
    return out.astype(bool)


def adjacent_counts(mine: np.ndarray) -> np.ndarray:
    m = np.ascontiguousarray(mine, dtype=np.uint8)
    out = np.zeros_like(m)
    lib().orc_adjacent_counts(m.shape[0], m.shape[1], m.ctypes.data, out.ctypes.data)
    return out


def flood_fill(revealed: np.ndarray, flags: np.ndarray, mine: np.ndarray, counts: np.ndarray,
               r: int, c: int) -> Tuple[np.ndarray, int]:
    """Returns (new revealed map, newly revealed count); env_numba.py:16-77."""
    rev = np.ascontiguousarray(revealed, dtype=np.uint8).copy()
    f = np.ascontiguousarray(flags, dtype=np.uint8)
    m = np.ascontiguousarray(mine, dtype=np.uint8)
    k = np.ascontiguousarray(counts, dtype=np.uint8)
    n = lib().orc_flood_fill(rev.shape[0], rev.shape[1], rev.ctypes.data, f.ctypes.data,
                             m.ctypes.data, k.ctypes.data, int(r), int(c))
    return rev.astype(bool), int(n)


def gae(rewards: np.ndarray, values: np.ndarray, dones: np.ndarray, last_values: np.ndarray,
        gamma: float = 0.995, lam: float = 0.95) -> Tuple[np.ndarray, np.ndarray]:
    """buffers.py:78-94 on [T,N] arrays; returns (advantages, returns) [T,N] f32."""
    r = np.ascontiguousarray(rewards, dtype=np.float32)
    v = np.ascontiguousarray(values, dtype=np.float32)
    d = np.ascontiguousarray(dones, dtype=np.uint8)
    lv = np.ascontiguousarray(last_values, dtype=np.float32)
    T, N = r.shape
    adv = np.zeros((T, N), dtype=np.float32)
    ret = np.zeros((T, N), dtype=np.float32)
    lib().orc_gae(T, N, r.ctypes.data, v.ctypes.data, d.ctypes.data, lv.ctypes.data,
                  float(np.float32(gamma)), float(np.float32(gamma * lam)), adv.ctypes.data, ret.ctypes.data)
    return adv, ret


class _EnvView:
    """Read-only stand-in for reference `vec.envs[i]` (env.py:68-75)."""

    def __init__(self, vec: "OracleVecEnv", i: int):
        self._v, self._i = vec, i
        self.cfg, self.H, self.W = vec.cfg, vec.H, vec.W

    @property
    def mine_mask(self): return self._v.mine[self._i].reshape(self.H, self.W).astype(bool)
    @property
    def revealed(self): return self._v.revealed[self._i].reshape(self.H, self.W).astype(bool)
    @property
    def flags(self): return self._v.flags[self._i].reshape(self.H, self.W).astype(bool)
    @property
    def adjacent_counts(self): return self._v.counts[self._i].reshape(self.H, self.W).copy()
    @property
    def first_click_done(self): return bool(self._v.first_click_done[self._i])
    @property
    def step_count(self): return int(self._v.step_count[self._i])


class OracleVecEnv:
    """CPU oracle with the reference VecMinesweeper call shape (env.py:379-517)."""

    def __init__(self, num_envs: int, cfg, seed: int = 0, env_id_base: int = 0,
                 nthreads: int = 1, aux_maps: bool = False, reuse_out: bool = False,
                 late_start: Optional[Tuple[int, float, int, int, int, int]] = None):
        assert num_envs > 0                                   # env.py:390
        self.cfg, self.num_envs, self.seed = cfg, int(num_envs), int(seed)
        self.H, self.W = int(cfg.H), int(cfg.W)
        self.HW = self.H * self.W
        self.env_id_base, self.nthreads, self.aux_maps = int(env_id_base), int(nthreads), aux_maps
        n, HW = self.num_envs, self.HW
        self.mine = np.zeros((n, HW), np.uint8)
        self.revealed = np.zeros((n, HW), np.uint8)
        self.flags = np.zeros((n, HW), np.uint8)
        self.counts = np.zeros((n, HW), np.uint8)
        self.first_click_done = np.zeros(n, np.int32)
        self.step_count = np.zeros(n, np.int32)
        self.last_new_reveals = np.zeros(n, np.int32)
        self.episode_idx = np.zeros(n, np.uint32)
        self._st = _State(*[a.ctypes.data for a in (
            self.mine, self.revealed, self.flags, self.counts, self.first_click_done,
            self.step_count, self.last_new_reveals, self.episode_idx)])
        self._ccfg = _ccfg(cfg, seed)
        self.envs = [_EnvView(self, i) for i in range(n)] if n <= 4096 else None
        self.mine_labels = self.mine_valid = None
        self.reuse_out, self._out = bool(reuse_out), None   # reuse_out: outputs are overwritten each call
        self.late_start = late_start    # (late_seed, prob, min_hidden, max_hidden, max_attempts, max_extra_steps)

    def action_space(self) -> int: return self.HW           # env.py:513-514
    def obs_channels(self) -> int: return OBS_CHANNELS      # env.py:516-517

    def _alloc_out(self):
        if self.reuse_out and self._out is not None:
            return self._out
        n, H, W = self.num_envs, self.H, self.W
        obs = np.empty((n, OBS_CHANNELS, H, W), np.float32)
        mask = np.empty((n, self.HW), np.uint8)
        lab = np.empty((n, H, W), np.float32) if self.aux_maps else None
        val = np.empty((n, H, W), np.uint8) if self.aux_maps else None
        self._out = (obs, mask, lab, val)
        return self._out

    def reset(self) -> Dict[str, np.ndarray]:
        obs, mask, lab, val = self._alloc_out()
        lib().orc_vec_reset(C.byref(self._ccfg), self.num_envs, C.byref(self._st),
                            _p(obs), _p(mask), _p(lab), _p(val), self.nthreads)
        self._late(None, obs, mask, lab, val)
        self.mine_labels, self.mine_valid = lab, (None if val is None else val.view(bool))
        return {"obs": obs, "action_mask": mask.view(bool)}

    def forced_subset(self) -> np.ndarray:
        """rules.analyze_forced_modules for every env: bool [n, HW] of "subset_reveal" cells."""
        out = np.zeros((self.num_envs, self.HW), np.uint8)
        lib().orc_forced_subset(C.byref(self._ccfg), self.num_envs, C.byref(self._st), out.ctypes.data)
        return out.astype(bool)

    def avoidability(self) -> Dict[str, np.ndarray]:
        """avoidability.analyze_avoidability for every env, in the array form documented in msw_oracle.h:
        {"safe": bool [n,HW], "comp_of_cell": i16 [n,HW], "comp_size": i16 [n,HW], "flags": u8 [n]}."""
        n, HW = self.num_envs, self.HW
        safe = np.zeros((n, HW), np.uint8)
        coc, cs = np.zeros((n, HW), np.int16), np.zeros((n, HW), np.int16)
        flags = np.zeros((n,), np.uint8)
        lib().orc_avoidability(C.byref(self._ccfg), n, C.byref(self._st), safe.ctypes.data, coc.ctypes.data,
                               cs.ctypes.data, flags.ctypes.data)
        return {"safe": safe.astype(bool), "comp_of_cell": coc, "comp_size": cs, "flags": flags}

    def _late(self, sel, obs, mask, lab, val):
        """env.py:406-414: late start on the envs just reset, then observe them again."""
        if not self.late_start:
            return
        seed, prob, lo, hi, attempts, extra = self.late_start
        lib().orc_late_start(C.byref(self._ccfg), self.num_envs, self.env_id_base, C.byref(self._st), _p(sel),
                             int(seed) & 0xFFFFFFFFFFFFFFFF, float(prob), int(lo), int(hi), int(attempts), int(extra))
        lib().orc_vec_encode(C.byref(self._ccfg), self.num_envs, C.byref(self._st), _p(obs), _p(mask), _p(lab),
                             _p(val), self.nthreads)

    def step(self, actions: np.ndarray, inject_mine: Optional[np.ndarray] = None,
             inject_sel: Optional[np.ndarray] = None, tensor_infos: bool = False
             ) -> Tuple[Dict[str, np.ndarray], np.ndarray, np.ndarray, Dict[str, Any]]:
        n = self.num_envs
        actions = np.asarray(actions)
        assert actions.shape == (n,)                          # env.py:480
        a64 = np.ascontiguousarray(actions, dtype=np.int64)
        obs, mask, lab, val = self._alloc_out()
        reward = np.empty(n, np.float32)
        done = np.empty(n, np.uint8)
        outcome = np.empty(n, np.int8)
        newrev = np.empty(n, np.int32)
        step = np.empty(n, np.int32)
        rcount = np.empty(n, np.int32)
        inj = None if inject_mine is None else np.ascontiguousarray(
            np.asarray(inject_mine).reshape(n, self.HW), dtype=np.uint8)
        sel = None if inject_sel is None else np.ascontiguousarray(inject_sel, dtype=np.uint8)
        out = _StepOut(_p(obs), _p(mask), _p(reward), _p(done), _p(outcome), _p(newrev), _p(step),
                       _p(rcount), _p(lab), _p(val))
        lib().orc_vec_step(C.byref(self._ccfg), n, self.env_id_base, a64.ctypes.data, _p(inj), _p(sel),
                           C.byref(self._st), C.byref(out), self.nthreads)
        self._late(done, obs, mask, lab, val)
        self.mine_labels, self.mine_valid = lab, (None if val is None else val.view(bool))
        dones = done.view(bool)
        if tensor_infos:
            infos: Dict[str, Any] = {"outcome_code": outcome, "last_new_reveals": newrev,
                                     "step": step, "revealed_count": rcount}
        else:                                                 # env.py:485-505
            names = (None, "win", "loss")
            infos = {
                "aux": [{"step": int(step[i]), "last_new_reveals": int(newrev[i]),
                         "revealed_frac": float(int(rcount[i]) / max(1, self.HW))} for i in range(n)],
                "outcome": [names[int(outcome[i])] for i in range(n)],
                "done": [bool(dones[i]) for i in range(n)],
            }
        return {"obs": obs, "action_mask": mask.view(bool)}, reward, dones, infos
