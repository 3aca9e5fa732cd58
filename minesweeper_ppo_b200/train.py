"""Thin PPO driver around the B200 rollout path (BASELINE.json config 5).

    python -m minesweeper_ppo_b200.train --config configs/medium_16x16x40.yaml --updates 20 [--envs-per-gpu N]
    torchrun --nproc-per-node 8 -m minesweeper_ppo_b200.train --config ... --updates 20

The PPO update itself (loss, AdamW, GradScaler, grad clipping) is NOT part of the accelerated path
("OUT OF SCOPE (kept PyTorch as-is; only an NCCL gradient-allreduce hook is added
around it)"); it is restated here from minesweeper/ppo.py:23-119 only so config 5 can run where the
reference tree is absent.  What this module adds is the multi-GPU shape of SURVEY section 8(e):
one process per GPU, each owning a contiguous range of global env ids (no communication on the
env / rollout / GAE path) and ONE NCCL all-reduce of the flattened gradient (~3.8 MB) per optimizer
step.  Not reproduced: checkpointing, quick-eval, CSV logging, aux-weight warm-up schedules
(train_rl.py:434-455, 515-541, 576-787).
"""
from __future__ import annotations

import argparse
import json
import os
import time
from dataclasses import dataclass, fields
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import fused_train as fused_train_mod
from .env import EnvConfig, VecMinesweeper
from .policy import build_model
from .rollout import RolloutCollector
from .shard import shard_range


@dataclass
class PPOConfig:                       # ppo.py:11-20
    clip_eps: float = 0.2
    clip_eps_v: float = 0.2
    vf_coef: float = 0.5
    ent_coef: float = 0.003
    aux_mine_weight: float = 0.0
    aux_mine_calib_weight: float = 0.0
    max_grad_norm: float = 0.5
    beta_l2: float = 0.0


@dataclass
class TrainConfig:                     # train_rl.py:82-107 (PPOTrainConfig)
    H: int = 8
    W: int = 8
    mine_count: int = 10
    guarantee_safe_neighborhood: bool = True
    num_envs: int = 256
    steps_per_env: int = 128
    mini_batches: int = 8
    ppo_epochs: int = 3
    gamma: float = 0.995
    gae_lambda: float = 0.95
    clip_eps: float = 0.2
    clip_eps_v: float = 0.2
    vf_coef: float = 0.5
    ent_coef: float = 0.003
    ent_coef_min: float = 0.003
    ent_decay_updates: int = 0
    lr: float = 3e-4
    max_grad_norm: float = 0.5
    aux_mine_weight: float = 0.0
    aux_mine_calib_weight: float = 0.0
    total_updates: int = 1000


def load_config(path: Optional[str]):
    """YAML with env / ppo / model / training sections (train_rl.py:110-143); training.rollout
    overrides ppo.num_envs / steps_per_env (train_rl.py:322-326)."""
    if path is None:
        return TrainConfig(), {}, {}, {}
    import yaml
    with open(path) as f:
        data = yaml.safe_load(f) or {}
    env_d, ppo_d, model_d = (data.get(k) or {} for k in ("env", "ppo", "model"))
    cfg = TrainConfig()
    for f_ in fields(TrainConfig):
        src = env_d if f_.name in ("H", "W", "mine_count", "guarantee_safe_neighborhood") else ppo_d
        if f_.name in src:
            setattr(cfg, f_.name, src[f_.name])
    extras = {k: v for k, v in data.items() if k not in ("env", "ppo", "model")}
    rollout = ((extras.get("training") or {}).get("rollout") or {})
    cfg.num_envs = int(rollout.get("num_envs", cfg.num_envs))
    cfg.steps_per_env = int(rollout.get("steps_per_env", cfg.steps_per_env))
    return cfg, env_d, model_d, extras


def ppo_loss(model, batch, cfg: PPOConfig, forward=None):
    """Clipped-ratio policy loss + clipped value loss - entropy bonus + auxiliary mine-belief
    BCE / Brier terms, as ppo.py:23-95 (fp16 autocast on CUDA, no advantage normalisation)."""
    on_cuda = batch.obs.is_cuda
    with torch.autocast(device_type="cuda", dtype=torch.float16, enabled=on_cuda):
        want_mine = cfg.aux_mine_weight > 0 or cfg.aux_mine_calib_weight > 0
        fwd = forward if forward is not None else model      # `forward`: e.g. the fused training forward
        if want_mine:
            logits, value, mine_logits = fwd(batch.obs, return_mine=True)
        else:
            (logits, value), mine_logits = fwd(batch.obs, return_mine=False), None
        fill = -1e4 if logits.dtype in (torch.float16, torch.bfloat16) else -1e9
        masked = logits.masked_fill(~batch.action_mask, fill)
        logp_all = F.log_softmax(masked, dim=-1)
        logp = logp_all.gather(1, batch.actions.unsqueeze(1)).squeeze(1)
        ratio = (logp - batch.old_logp).exp()
        surrogate = torch.min(ratio * batch.advantages,
                              ratio.clamp(1 - cfg.clip_eps, 1 + cfg.clip_eps) * batch.advantages)
        policy_loss = -surrogate.mean()
        v = value.view(-1)
        v_clip = batch.values + (v - batch.values).clamp(-cfg.clip_eps_v, cfg.clip_eps_v)
        value_loss = 0.5 * torch.max((v - batch.returns).pow(2), (v_clip - batch.returns).pow(2)).mean()
        entropy = -(torch.softmax(masked, -1) * logp_all).sum(-1).mean()
        loss = policy_loss + cfg.vf_coef * value_loss - cfg.ent_coef * entropy
        stats = {"policy_loss": policy_loss, "value_loss": value_loss, "entropy": entropy}
        if want_mine and mine_logits is not None and hasattr(batch, "mine_labels"):
            ml = mine_logits.squeeze(1)
            labels = batch.mine_labels
            valid = getattr(batch, "mine_valid", None)
            if valid is None:
                valid = torch.ones_like(labels, dtype=torch.bool)
            zl, yl = ml[valid], labels[valid]
            if yl.numel() > 0:
                if cfg.aux_mine_weight > 0:                     # per-minibatch pos_weight, ppo.py:67-71
                    pos = yl.sum()
                    pw = float(((yl.numel() - pos) + 1e-6) / (pos + 1e-6))
                    bce = F.binary_cross_entropy_with_logits(zl, yl, pos_weight=torch.full((), pw, dtype=zl.dtype, device=zl.device))
                    loss = loss + cfg.aux_mine_weight * bce
                    stats["aux_bce"] = bce
                if cfg.aux_mine_calib_weight > 0:
                    brier = (torch.sigmoid(zl) - yl).pow(2).mean()
                    loss = loss + cfg.aux_mine_calib_weight * brier
                    stats["aux_calib"] = brier
        if cfg.beta_l2 > 0 and hasattr(model, "beta_regularizer"):
            loss = loss + cfg.beta_l2 * model.beta_regularizer()
    stats["loss"] = loss
    return loss, stats


class FlatGradAllReduce:
    """Data parallelism over env shards = ONE NCCL all-reduce per optimizer step: the gradients
    that exist after backward (parameters outside the graph keep grad=None, exactly as in the
    reference, so AdamW skips them) are packed into one flat bucket (~3.8 MB for the medium
    model), summed over ranks, averaged and unpacked."""

    def __init__(self, model: torch.nn.Module, time_it: bool = False):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.numel = sum(p.numel() for p in self.params)
        # The bucket layout is FIXED at the first sync (the parameters that have a gradient then); every later
        # step packs exactly that set, a rank whose minibatch produced no gradient for one of them (e.g. the
        # mine head when no label cell is valid) contributes zeros, and a parameter that turns up with a
        # gradient outside the set is an error -- so all ranks always reduce buffers of the same size.
        self._active = None
        self._flat = None
        self.time_it = bool(time_it)
        self._events = []          # (start, end) CUDA events around each all-reduce when time_it

    def sync(self) -> None:
        if self.world <= 1:
            return
        if self._active is None:
            self._active = [p for p in self.params if p.grad is not None]
            sizes = torch.tensor([len(self._active), sum(p.numel() for p in self._active)], dtype=torch.int64,
                                 device=self._active[0].device)
            lo, hi = sizes.clone(), sizes.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                raise RuntimeError("FlatGradAllReduce: ranks disagree on which parameters have gradients "
                                   f"(min {lo.tolist()}, max {hi.tolist()})")
            self._flat = torch.empty(int(sizes[1]), dtype=self._active[0].grad.dtype, device=sizes.device)
            self._ids = {id(p) for p in self._active}
        for p in self.params:
            if p.grad is not None and id(p) not in self._ids:
                raise RuntimeError("FlatGradAllReduce: a parameter outside the bucket fixed at the first step has a gradient")
        flat, off = self._flat, 0
        for p in self._active:
            n = p.numel()
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if self.time_it:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if self.time_it:
            e1.record()
            self._events.append((e0, e1))
        flat.div_(self.world)
        gs = [p.grad for p in self._active]
        torch._foreach_copy_(gs, [c.view_as(g) for c, g in zip(flat.split([g.numel() for g in gs]), gs)])

    def allreduce_ms(self, skip: int = 0):
        """Mean / max device time of one all-reduce (after `skip` warm-up calls); needs time_it=True."""
        ev = self._events[skip:]
        if not ev:
            return None
        ms = [a.elapsed_time(b) for a, b in ev]
        return {"mean_ms": sum(ms) / len(ms), "max_ms": max(ms), "calls": len(ms)}


def ppo_update(model, optimizer, batch, cfg: PPOConfig, scaler=None, grads: Optional[FlatGradAllReduce] = None,
               want_stats: bool = True, forward=None) -> Dict[str, float]:
    """One optimizer step (ppo.py:96-119) with the gradient all-reduce between backward and
    unscale/clip."""
    loss, stats = ppo_loss(model, batch, cfg, forward=forward)
    optimizer.zero_grad(set_to_none=True)
    if scaler is not None and batch.obs.is_cuda:
        scaler.scale(loss).backward()
        if grads is not None:
            grads.sync()
        scaler.unscale_(optimizer)
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.max_grad_norm)
        scaler.step(optimizer)
        scaler.update()
    else:
        loss.backward()
        if grads is not None:
            grads.sync()
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.max_grad_norm)
        optimizer.step()
    return {k: float(v) for k, v in stats.items()} if want_stats else {}


def train(config: Optional[str], updates: int, envs_per_gpu: Optional[int], steps: Optional[int], seed: int = 0,
          log: Callable[[str], None] = print, fused_train: bool = True, time_allreduce: bool = False) -> Dict[str, float]:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    own_group = world > 1 and not dist.is_initialized()
    if own_group:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg, env_d, model_d, extras = load_config(config)
    n_local = int(envs_per_gpu or cfg.num_envs)
    T = int(steps or cfg.steps_per_env)
    base, _ = shard_range(n_local * world, rank, world)
    env_kwargs = {k: v for k, v in env_d.items() if k != "include_frontier_channel"}      # train_rl.py:348
    env_cfg = EnvConfig(**env_kwargs) if env_kwargs else EnvConfig(H=cfg.H, W=cfg.W, mine_count=cfg.mine_count)
    vec = VecMinesweeper(n_local, env_cfg, seed=seed, api="torch", env_id_base=base)

    torch.manual_seed(seed)                       # identical initial weights on every rank
    mcfg = dict(model_d)
    name = mcfg.pop("name", "cnn")
    model = build_model(name, obs_shape=(vec.obs_channels(), vec.H, vec.W), model_cfg=mcfg).to(dev)
    torch.manual_seed(seed + 1 + rank)            # dropout / minibatch permutations differ per rank
    opt = torch.optim.AdamW(model.parameters(), lr=cfg.lr)
    scaler = torch.amp.GradScaler("cuda")
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=cfg.total_updates)
    pcfg = PPOConfig(clip_eps=cfg.clip_eps, clip_eps_v=cfg.clip_eps_v, vf_coef=cfg.vf_coef, ent_coef=cfg.ent_coef,
                     aux_mine_weight=cfg.aux_mine_weight, aux_mine_calib_weight=cfg.aux_mine_calib_weight,
                     max_grad_norm=cfg.max_grad_norm, beta_l2=float((extras.get("training") or {}).get("beta_l2", 0.0)))
    need_aux = pcfg.aux_mine_weight > 0 or pcfg.aux_mine_calib_weight > 0
    collector = RolloutCollector(vec, T, aux_maps=need_aux, sample_seed=seed, graph=RolloutCollector.can_graph(model))
    grads = FlatGradAllReduce(model, time_it=time_allreduce)
    mb = (n_local * T) // cfg.mini_batches
    opt_step = [0]
    train_fwd = None
    if fused_train and fused_train_mod.supports(model):      # fused GroupNorm forward + backward
        def train_fwd(obs, return_mine=False):
            opt_step[0] += 1
            return fused_train_mod.fused_train_forward(model, obs, return_mine, seed=seed + 7919 * rank, step=opt_step[0])

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    timed_from = 1 if updates > 1 else 0           # first update pays cuDNN autotune / allocator warm-up
    t_roll = t_upd = 0.0
    last = {}
    for u in range(updates):
        if cfg.ent_decay_updates > 0:              # train_rl.py:515-523
            frac = min(1.0, u / max(1, int(cfg.ent_decay_updates)))
            pcfg.ent_coef = float(cfg.ent_coef + (cfg.ent_coef_min - cfg.ent_coef) * frac)
        sync(); t0 = time.perf_counter()
        buf, aux = collector.collect(model)
        buf.compute_gae(aux["last_values"], gamma=cfg.gamma, lam=cfg.gae_lambda)
        sync(); t1 = time.perf_counter()
        for _ in range(cfg.ppo_epochs):
            for batch in buf.get_minibatches(mb):
                last = ppo_update(model, opt, batch, pcfg, scaler, grads, want_stats=False, forward=train_fwd)
        sched.step()
        sync(); t2 = time.perf_counter()
        if u >= timed_from:
            t_roll += t1 - t0
            t_upd += t2 - t1
        if rank == 0:
            log(f"update {u}: rollout {1e3 * (t1 - t0):.1f} ms, ppo {1e3 * (t2 - t1):.1f} ms, "
                f"episodes {int(buf.dones.sum())}, mean reward {float(buf.rewards.mean()):.4f}")
    n_timed = max(1, updates - timed_from)
    total = t_roll + t_upd
    # replicas must still agree (same init, averaged gradients)
    check = torch.stack([p.detach().float().sum() for p in model.parameters()]).sum()
    spread = 0.0
    if world > 1:
        lo, hi = check.clone(), check.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = float(hi - lo)
    result = {
        "workload": "C5", "n_gpus": world, "envs_per_gpu": n_local, "steps_per_env": T, "updates_timed": n_timed,
        "updates_per_s": n_timed / total, "frames_per_s": world * n_local * T * n_timed / total,
        "rollout_frames_per_s": world * n_local * T * n_timed / t_roll,
        "ms_rollout_gae": 1e3 * t_roll / n_timed, "ms_ppo_epochs": 1e3 * t_upd / n_timed,
        "optimizer_steps_per_update": cfg.ppo_epochs * cfg.mini_batches,
        "grad_allreduce_bytes": grads.numel * 4, "replica_param_checksum_spread": spread,
        "training_forward": ("fused GroupNorm fwd+bwd (msw_gn_act / msw_gn_act_bwd), trunk conv fwd + dgrad on tcgen05 (msw_conv3x3)"
                             if train_fwd else "eager module"),
    }
    if time_allreduce and world > 1:
        torch.cuda.synchronize()
        result["grad_allreduce"] = grads.allreduce_ms(skip=cfg.ppo_epochs * cfg.mini_batches * timed_from)
    if own_group:
        dist.destroy_process_group()
    return result if rank == 0 else {}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=None)
    ap.add_argument("--updates", type=int, default=5)
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--eager-train", action="store_true", help="use the unchanged module for the PPO update")
    a = ap.parse_args()
    r = train(a.config, a.updates, a.envs_per_gpu, a.steps, a.seed, log=lambda s: print(s, flush=True),
              fused_train=not a.eager_train)
    if r:
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
