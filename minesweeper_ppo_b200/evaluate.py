"""Greedy vectorised evaluation through the reference-shaped (NumPy) env API -- BASELINE.json
configs[0] (C1).

This is the *consumer pattern* of `eval.evaluate_vec` (eval.py:265-511) on the drop-in boundary:
`VecMinesweeper(num_envs, env_cfg, seed)` built by the callee, NumPy obs/mask each step, greedy argmax
of the masked logits, per-env reads of `vec.envs[i].first_click_done / revealed / flags / mine_mask`
for the belief statistics (eval.py:350-360), `vec.step(actions)` and the `infos["aux"] / ["outcome"]`
lists (eval.py:405-416), with the reference's episode accounting (only the first finished episode
per env and batch counts, eval.py:411-428).  The solver / avoidability analytics of the reference
(rules.py, avoidability.py) are out of scope and not reproduced, so only the metrics that do not need
them are returned.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from .env import EnvConfig, VecMinesweeper
from .rules import analyze_forced_modules


def _auroc(labels: np.ndarray, scores: np.ndarray) -> float:
    """Rank-based AUROC (ties get average ranks)."""
    pos = labels > 0.5
    n_pos, n_neg = int(pos.sum()), int((~pos).sum())
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    order = np.argsort(scores, kind="mergesort")
    ranks = np.empty(len(scores), dtype=np.float64)
    sorted_scores = scores[order]
    i = 0
    while i < len(scores):
        j = i
        while j + 1 < len(scores) and sorted_scores[j + 1] == sorted_scores[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return float((ranks[pos].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


@torch.no_grad()
def evaluate_vec(model: torch.nn.Module, env_cfg: EnvConfig, episodes: int = 1000, seed: int = 0,
                 num_envs: int = 256, max_steps_per_episode: int = 512) -> Dict[str, float]:
    device = next(model.parameters()).device
    was_training = model.training
    model.eval()                                                          # eval.py:278-279
    vec = VecMinesweeper(num_envs=num_envs, cfg=env_cfg, seed=seed)       # reference call shape (NumPy API)
    batch = vec.reset()
    HW = env_cfg.H * env_cfg.W
    remaining, wins, total_steps, total_progress, invalids = episodes, 0, 0, 0.0, 0
    probs, labels = [], []
    forced_steps = forced_correct = guess_attempts = guess_success = 0
    while remaining > 0:
        batch_size = min(num_envs, remaining)
        finished = 0
        counted = np.zeros(num_envs, dtype=bool)
        step_counters = np.zeros(num_envs, dtype=np.int32)
        while finished < batch_size:
            obs = torch.from_numpy(batch["obs"]).to(device=device, dtype=torch.float32)
            mask = torch.from_numpy(batch["action_mask"]).to(device=device, dtype=torch.bool)
            empty = ~mask.any(dim=1)
            if empty.any():
                mask[empty] = True
            logits, _, mine_logits = model(obs, return_mine=True)         # fp32, no autocast (eval.py:334)
            assert logits.shape[1] == mask.shape[1] == vec.envs[0].action_space      # eval.py:22-27
            actions = logits.masked_fill(~mask, -1e9).argmax(dim=-1).cpu().numpy().astype(np.int32)
            picked = mask.cpu().numpy()[np.arange(num_envs), actions]
            invalids += int((~picked).sum())
            mine_prob = torch.sigmoid(mine_logits).cpu().numpy()
            for idx, env in enumerate(vec.envs):                          # eval.py:350-360
                if counted[idx] or not env.first_click_done:
                    continue
                unknown = (~env.revealed) & (~env.flags)
                if unknown.any():
                    probs.append(mine_prob[idx, 0][unknown])
                    labels.append(env.mine_mask[unknown].astype(np.float32))
                # forced reveals by the subset rule (eval.py:362-379); one batched kernel per step
                cell = int(actions[idx])
                safe = not env.mine_mask[cell // env.W, cell % env.W]
                if cell in analyze_forced_modules(env)["subset_reveal"]:
                    forced_steps += 1; forced_correct += safe
                else:
                    guess_attempts += 1; guess_success += safe
            batch, rewards, dones, infos = vec.step(actions)
            step_counters += 1
            for i in range(num_envs):                                     # eval.py:405-428
                if not counted[i]:
                    total_progress += int(infos["aux"][i].get("last_new_reveals", 0)) / float(HW)
                if not counted[i] and dones[i]:
                    wins += infos["outcome"][i] == "win"
                    total_steps += int(step_counters[i]); step_counters[i] = 0
                    counted[i] = True; finished += 1
                if not counted[i] and 0 < max_steps_per_episode <= step_counters[i]:
                    total_steps += int(step_counters[i]); step_counters[i] = 0
                    counted[i] = True; finished += 1
        remaining -= batch_size
    if was_training:
        model.train()
    p = np.concatenate(probs) if probs else np.zeros(0)
    y = np.concatenate(labels) if labels else np.zeros(0)
    return {
        "win_rate": wins / max(1, episodes), "avg_steps": total_steps / max(1, episodes),
        "avg_progress": total_progress / max(1, episodes), "invalid_rate": invalids / max(1, total_steps),
        "belief_auroc": _auroc(y, p) if len(p) else float("nan"),
        "forced_step_rate": forced_steps / max(1, forced_steps + guess_attempts),
        "forced_step_accuracy": forced_correct / max(1, forced_steps),
        "guess_success_rate": guess_success / max(1, guess_attempts),
        "wins": float(wins), "episodes": float(episodes),
    }
