"""Rollout storage + GAE with the reference's `RolloutBuffer` API (minesweeper/buffers.py:9-116).

Same flat, time-major layout (`index = t*N + i`, buffers.py:51-52), same field names and
dtypes (buffers.py:24-34), so `ppo_update` (ppo.py:23-119) consumes it unchanged.  Two things
differ from the reference implementation:

  * `compute_gae` is ONE CUDA kernel (msw_gae) instead of a T-step Python loop of ~8 torch
    ops (buffers.py:87-92); results are bit-identical.
  * `slot(t)` hands out views of step t so the env kernel writes obs / mask / reward / done /
    aux maps straight into the buffer instead of going through `add`'s seven slice copies
    (buffers.py:53-59).  `add` itself is kept with identical semantics.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Tuple

import numpy as np
import torch

import ctypes as C

from . import _lib
from .env import StepOut, VecMinesweeper


class RolloutBuffer:
    def __init__(
        self,
        num_envs: int,
        steps: int,
        obs_shape: Tuple[int, int, int],
        action_dim: int,
        device: torch.device,
        aux_maps: bool = False,
    ):
        self.num_envs = num_envs
        self.steps = steps
        self.device = torch.device(device)

        B = num_envs * steps
        C, H, W = obs_shape
        dev = self.device
        self.obs = torch.zeros((B, C, H, W), dtype=torch.float32, device=dev)
        self.action_mask = torch.zeros((B, action_dim), dtype=torch.bool, device=dev)
        self.actions = torch.zeros((B,), dtype=torch.long, device=dev)
        self.logp = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((B,), dtype=torch.bool, device=dev)
        self.values = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.advantages = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.returns = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.mine_labels: Optional[torch.Tensor] = None
        self.mine_valid: Optional[torch.Tensor] = None
        if aux_maps:                                   # the reference allocates these lazily (buffers.py:60-75)
            self.mine_labels = torch.zeros((B, H, W), dtype=torch.float32, device=dev)
            self.mine_valid = torch.zeros((B, H, W), dtype=torch.bool, device=dev)

        self._t = 0

    # ---------------------------------------------------------------- direct-write protocol
    def slot(self, t: int) -> StepOut:
        """Views of time slot t for the env kernel.  Time alignment follows train_rl.py:196-265:
        slot t holds obs_t / mask_t / labels_t / valid_t (the state BEFORE action a_t) and
        reward_t / done_t (the result OF a_t) -- so pass `obs/action_mask/mine_*` of slot(t+1)
        together with `rewards/dones` of slot(t) to `vec.step` (see rollout.collect_rollout)."""
        n = self.num_envs
        s, e = t * n, (t + 1) * n
        return StepOut(
            obs=self.obs[s:e], action_mask=self.action_mask[s:e], rewards=self.rewards[s:e], dones=self.dones[s:e],
            mine_labels=None if self.mine_labels is None else self.mine_labels[s:e],
            mine_valid=None if self.mine_valid is None else self.mine_valid[s:e],
        )

    # ---------------------------------------------------------------- reference API
    def add(
        self,
        obs: torch.Tensor,
        action_mask: torch.Tensor,
        actions: torch.Tensor,
        logp: torch.Tensor,
        rewards: torch.Tensor,
        dones: torch.Tensor,
        values: torch.Tensor,
        mine_labels: Optional[torch.Tensor] = None,
        mine_valid: Optional[torch.Tensor] = None,
    ) -> None:
        """Same contract as buffers.py:38-76: copy one step into slot `_t`, then advance."""
        n = obs.shape[0]
        rows = slice(self._t * n, (self._t + 1) * n)
        for name, src in (("obs", obs), ("action_mask", action_mask), ("actions", actions), ("logp", logp),
                          ("rewards", rewards), ("dones", dones), ("values", values)):
            getattr(self, name)[rows] = src
        if mine_labels is not None:                      # aux maps appear lazily, labels gate valid
            self._ensure_aux(obs.shape[-2], obs.shape[-1], with_valid=mine_valid is not None)
            self.mine_labels[rows] = mine_labels
            if mine_valid is not None:
                self.mine_valid[rows] = mine_valid
        self._t += 1

    def _ensure_aux(self, H: int, W: int, with_valid: bool) -> None:
        B = self.num_envs * self.steps
        if self.mine_labels is None:
            self.mine_labels = torch.zeros((B, H, W), dtype=torch.float32, device=self.device)
        if with_valid and self.mine_valid is None:
            self.mine_valid = torch.zeros((B, H, W), dtype=torch.bool, device=self.device)

    def compute_gae(self, last_values: torch.Tensor, gamma: float = 0.995, lam: float = 0.95) -> None:
        """buffers.py:78-94 as one kernel launch on the current stream."""
        if self.device.type != "cuda":
            raise RuntimeError("RolloutBuffer.compute_gae runs on CUDA only (no CPU fallback in this package)")
        L = _lib.load()
        N, T = self.num_envs, self.steps
        lv = last_values.detach().reshape(N).to(self.device)
        prescaled = 0
        if lv.dtype != torch.float32:
            # Reference subtlety (buffers.py:88-90 with the fp16 bootstrap value of
            # train_rl.py:272-277): `gamma * next_value` is evaluated in the tensor's own dtype
            # at t = T-1.  Let torch do exactly that multiply, then hand the kernel the product.
            lv = (gamma * lv).float()
            prescaled = 1
        lv = lv.contiguous()
        if self.advantages.shape != (T * N,) or not self.advantages.is_contiguous():
            self.advantages = torch.empty((T * N,), dtype=torch.float32, device=self.device)
        if self.returns.shape != (T * N,) or not self.returns.is_contiguous():
            self.returns = torch.empty((T * N,), dtype=torch.float32, device=self.device)
        gamma_f32 = float(np.float32(gamma))
        gamma_lam_f32 = float(np.float32(gamma * lam))          # product in float64 first (buffers.py:91)
        with torch.cuda.device(self.device):
            rc = L.msw_gae(self.rewards.data_ptr(), self.values.data_ptr(), self.dones.data_ptr(), lv.data_ptr(),
                           self.advantages.data_ptr(), self.returns.data_ptr(), T, N, gamma_f32, gamma_lam_f32,
                           prescaled, torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "msw_gae")

    def get_minibatches(self, batch_size: int) -> Iterator[Dict[str, torch.Tensor]]:
        """Same contract as buffers.py:96-116: one random permutation of all T*N rows, yielded in
        chunks as attribute-style batches (SURVEY 8f row f2 is the fused gather)."""
        total = self.obs.shape[0]
        order = torch.randperm(total, device=self.device)
        fields = [("obs", self.obs), ("action_mask", self.action_mask), ("actions", self.actions),
                  ("old_logp", self.logp), ("rewards", self.rewards), ("dones", self.dones),
                  ("values", self.values), ("advantages", self.advantages), ("returns", self.returns)]
        if self.mine_labels is not None:
            fields.append(("mine_labels", self.mine_labels))
            if self.mine_valid is not None:
                fields.append(("mine_valid", self.mine_valid))
        for start in range(0, total, batch_size):
            rows = order[start: start + batch_size]
            yield type("Batch", (), {name: tensor[rows] for name, tensor in fields})


class CompactRolloutBuffer(RolloutBuffer):
    """RolloutBuffer that stores each transition's OBSERVATION as the bitboards it was encoded from
    (SURVEY section 8 row f2): mines + revealed (+ flags) + first_click_done = 65 B per transition
    at 16x16 instead of 11.8 KB of obs / action_mask / mine_labels / mine_valid, and re-encodes the
    rows of a minibatch on the fly (msw_gather_encode) inside `get_minibatches`.  Batches are
    bit-identical to the dense buffer's for the same permutation; the dense `obs` / `action_mask` /
    `mine_*` attributes do not exist."""

    def __init__(self, vec: VecMinesweeper, steps: int, aux_maps: bool = False):
        n, dev = vec.num_envs, vec.device
        self.num_envs, self.steps, self.device = n, int(steps), dev
        self.vec, self.aux_maps = vec, bool(aux_maps)
        B = n * self.steps
        self.snap_mines = torch.zeros((B, vec.wpb), dtype=torch.int32, device=dev)
        self.snap_revealed = torch.zeros((B, vec.wpb), dtype=torch.int32, device=dev)
        self.snap_flags: Optional[torch.Tensor] = None
        self.snap_first = torch.zeros((B,), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((B,), dtype=torch.long, device=dev)
        self.logp = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((B,), dtype=torch.bool, device=dev)
        self.values = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.advantages = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.returns = torch.zeros((B,), dtype=torch.float32, device=dev)
        self.obs = self.action_mask = self.mine_labels = self.mine_valid = None
        self._t = 0

    def snapshot(self, t: int) -> None:
        """Record the env's CURRENT state as the observation of slot t (call right after the
        reset / step whose observation slot t holds)."""
        n = self.num_envs
        rows = slice(t * n, (t + 1) * n)
        st = self.vec.state_tensors
        self.snap_mines[rows] = st["mines"]
        self.snap_revealed[rows] = st["revealed"]
        self.snap_first[rows] = st["meta"][:, 0].to(torch.uint8)
        if st["flags"] is not None:
            if self.snap_flags is None:
                self.snap_flags = torch.zeros_like(self.snap_mines)
            self.snap_flags[rows] = st["flags"]

    def slot(self, t: int) -> StepOut:
        raise RuntimeError("CompactRolloutBuffer has no dense observation slots; use snapshot(t)")

    def add(self, *a, **k) -> None:
        raise RuntimeError("CompactRolloutBuffer is filled by RolloutCollector(compact=True)")

    def gather_obs(self, rows: torch.Tensor) -> StepOut:
        """Dense obs / mask (/ aux maps) of the given transition rows, re-encoded on device."""
        L = _lib.load()
        v = self.vec
        m = int(rows.numel())
        rows = rows.to(device=self.device, dtype=torch.int64).contiguous()
        out = StepOut(
            obs=torch.empty((m, v.obs_channels(), v.H, v.W), dtype=torch.float32, device=self.device),
            action_mask=torch.empty((m, v.HW), dtype=torch.bool, device=self.device),
            mine_labels=torch.empty((m, v.H, v.W), dtype=torch.float32, device=self.device) if self.aux_maps else None,
            mine_valid=torch.empty((m, v.H, v.W), dtype=torch.bool, device=self.device) if self.aux_maps else None)
        enc = _lib.EncodeOut(out.obs.data_ptr(), out.action_mask.data_ptr(),
                             None if out.mine_labels is None else out.mine_labels.data_ptr(),
                             None if out.mine_valid is None else out.mine_valid.data_ptr())
        with torch.cuda.device(self.device):
            rc = L.msw_gather_encode(C.byref(v._desc), self.snap_mines.data_ptr(), self.snap_revealed.data_ptr(),
                                     None if self.snap_flags is None else self.snap_flags.data_ptr(),
                                     self.snap_first.data_ptr(), self.snap_mines.shape[0], rows.data_ptr(), m,
                                     C.byref(enc), torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "msw_gather_encode")
        return out

    def get_minibatches(self, batch_size: int) -> Iterator[Dict[str, torch.Tensor]]:
        total = self.num_envs * self.steps
        order = torch.randperm(total, device=self.device)
        small = [("actions", self.actions), ("old_logp", self.logp), ("rewards", self.rewards), ("dones", self.dones),
                 ("values", self.values), ("advantages", self.advantages), ("returns", self.returns)]
        for start in range(0, total, batch_size):
            rows = order[start: start + batch_size]
            dense = self.gather_obs(rows)
            batch = {"obs": dense.obs, "action_mask": dense.action_mask}
            batch.update({name: tensor[rows] for name, tensor in small})
            if self.aux_maps:
                batch["mine_labels"], batch["mine_valid"] = dense.mine_labels, dense.mine_valid
            yield type("Batch", (), batch)
