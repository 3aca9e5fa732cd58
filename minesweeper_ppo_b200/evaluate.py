"""Greedy vectorised evaluation through the reference-shaped (NumPy) env API -- BASELINE.json
configs[0] (C1).

This is `eval.evaluate_vec` (eval.py:265-511) restated on the drop-in boundary: the callee builds
`VecMinesweeper(num_envs, env_cfg, seed)`, reads NumPy obs/mask each step, plays the greedy argmax of
the masked logits, reads `vec.envs[i].first_click_done / revealed / flags / mine_mask` for the belief
statistics (eval.py:350-360), calls `analyze_forced_modules(env)` and `analyze_avoidability(env, cell)`
per live env (eval.py:362-398) -- here two batched kernel launches per step behind the reference's
call shape -- then `vec.step(actions)` and the `infos["aux"] / ["outcome"]` lists (eval.py:405-416),
with the reference's episode accounting (only the first finished episode per env and batch counts,
eval.py:411-428).  Returns the reference's metric dict, key for key.  `tests/golden/c1_eval.npz` holds
the dict the unmodified reference produced for a scripted policy; `test_eval_matches_reference_metrics`
replays the same mine layouts through this function on the CUDA env and compares.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from .avoidability import analyze_avoidability
from .env import EnvConfig, VecMinesweeper
from .rules import analyze_forced_modules


def _auroc(labels: np.ndarray, scores: np.ndarray) -> float:
    """Rank AUROC as the reference computes it (eval.py:54-66): plain argsort ranks, ties not averaged."""
    labels, scores = labels.reshape(-1), scores.reshape(-1)
    n_pos, n_neg = float((labels == 1).sum()), float((labels == 0).sum())
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    order = scores.argsort()
    ranks = np.empty(len(scores), dtype=np.float64)
    ranks[order] = np.arange(1, len(scores) + 1, dtype=np.float64)
    return float((ranks[labels == 1].sum() - n_pos * (n_pos + 1.0) / 2.0) / (n_pos * n_neg))


def _ece(probs: np.ndarray, labels: np.ndarray, bins: int = 15) -> float:
    """Expected calibration error over equal-width bins, last bin closed (eval.py:69-90)."""
    probs, labels = probs.reshape(-1), labels.reshape(-1)
    if probs.shape[0] == 0:
        return float("nan")
    edges = np.linspace(0.0, 1.0, bins + 1)
    total = 0.0
    for k in range(bins):
        sel = (probs >= edges[k]) & ((probs <= edges[k + 1]) if k == bins - 1 else (probs < edges[k + 1]))
        cnt = sel.sum()
        if cnt:
            total += (cnt / probs.shape[0]) * abs(labels[sel].mean() - probs[sel].mean())
    return float(total)


def _wilson(successes: int, total: int, z: float = 1.96):
    """95 % Wilson score interval (eval.py:447-458)."""
    if total <= 0:
        return float("nan"), float("nan")
    phat = successes / float(total)
    denom = 1.0 + (z * z) / total
    center = phat + (z * z) / (2.0 * total)
    rad = z * np.sqrt((phat * (1.0 - phat) / total) + (z * z) / (4.0 * total * total))
    return float((center - rad) / denom), float((center + rad) / denom)


def _ratio(num: float, den: float) -> float:
    return num / float(den) if den > 0 else float("nan")


@torch.no_grad()
def evaluate_vec(model: torch.nn.Module, env_cfg: EnvConfig, episodes: int = 1000, seed: int = 0,
                 num_envs: int = 256, max_steps_per_episode: int = 512,
                 vec_factory: Optional[Callable[..., VecMinesweeper]] = None) -> Dict[str, float]:
    """`vec_factory(num_envs=, cfg=, seed=)` lets a harness substitute the env the reference would build
    itself (the parity test injects the reference's mine layouts that way)."""
    device = next(model.parameters()).device
    was_training = model.training
    model.eval()                                                          # eval.py:278-279
    vec = (vec_factory or VecMinesweeper)(num_envs=num_envs, cfg=env_cfg, seed=seed)   # NumPy API
    batch = vec.reset()
    HW = env_cfg.H * env_cfg.W
    remaining, wins, total_steps, total_progress, invalids = episodes, 0, 0, 0.0, 0
    probs, labels = [], []
    forced_steps = forced_correct = guess_attempts = guess_success = 0
    reveal_total = forced_guess_total = forced_guess_success = forced_guess_episodes = 0
    safe_option_total = safe_option_hits = safe_option_misses = safe_cells = 0
    component_sizes: List[int] = []
    chosen_sizes: List[int] = []
    while remaining > 0:
        batch_size = min(num_envs, remaining)
        finished = 0
        counted = np.zeros(num_envs, dtype=bool)
        step_counters = np.zeros(num_envs, dtype=np.int32)
        ep_unavoidable = np.zeros(num_envs, dtype=bool)
        while finished < batch_size:
            obs = torch.from_numpy(batch["obs"]).to(device=device, dtype=torch.float32)
            mask = torch.from_numpy(batch["action_mask"]).to(device=device, dtype=torch.bool)
            empty = ~mask.any(dim=1)
            if empty.any():
                mask[empty] = True
            logits, _, mine_logits = model(obs, return_mine=True)         # fp32, no autocast (eval.py:333)
            assert logits.shape[1] == mask.shape[1] == vec.envs[0].action_space      # eval.py:22-27
            actions = logits.masked_fill(~mask, -1e9).argmax(dim=-1).cpu().numpy().astype(np.int32)
            picked = mask.cpu().numpy()[np.arange(num_envs), actions]
            invalids += int((~picked).sum())
            mine_prob = torch.sigmoid(mine_logits).cpu().numpy() if mine_logits is not None else None
            for idx, env in enumerate(vec.envs):                          # eval.py:350-398
                if counted[idx] or idx >= batch_size:
                    continue
                cell = int(actions[idx])
                safe = not env.mine_mask[cell // env.W, cell % env.W]
                if mine_prob is not None:
                    unknown = (~env.revealed) & (~env.flags)
                    if unknown.any():
                        probs.append(mine_prob[idx, 0][unknown].reshape(-1))
                        labels.append(env.mine_mask[unknown].astype(np.float32).reshape(-1))
                # forced reveals by the subset rule (eval.py:362-379); one batched kernel per step
                if cell in analyze_forced_modules(env)["subset_reveal"]:
                    forced_steps += 1; forced_correct += safe
                else:
                    guess_attempts += 1; guess_success += safe
                if env.first_click_done:                                  # eval.py:381-398, batched likewise
                    res = analyze_avoidability(env, cell)
                    component_sizes.extend(res.component_sizes)
                    if res.chosen_component_size is not None:
                        chosen_sizes.append(res.chosen_component_size)
                    reveal_total += 1
                    if res.avoidable:
                        safe_option_total += 1
                        safe_cells += res.count_forced_safe_cells
                        if res.chosen_is_forced_safe:
                            safe_option_hits += 1
                        else:
                            safe_option_misses += 1
                    else:
                        forced_guess_total += 1
                        ep_unavoidable[idx] = True
                        forced_guess_success += safe
            batch, rewards, dones, infos = vec.step(actions)
            step_counters += 1
            for i in range(num_envs):                                     # eval.py:405-428
                if not counted[i]:
                    total_progress += int(infos["aux"][i].get("last_new_reveals", 0)) / float(HW)
                if not counted[i] and dones[i]:
                    wins += infos["outcome"][i] == "win"
                    total_steps += int(step_counters[i]); step_counters[i] = 0
                    counted[i] = True; finished += 1
                    forced_guess_episodes += bool(ep_unavoidable[i])
                if not counted[i] and 0 < max_steps_per_episode <= step_counters[i]:
                    total_steps += int(step_counters[i]); step_counters[i] = 0
                    counted[i] = True; finished += 1
        remaining -= batch_size
    if was_training:
        model.train()
    p = np.concatenate(probs) if probs else np.zeros(0)
    y = np.concatenate(labels) if labels else np.zeros(0)
    ci_low, ci_high = _wilson(wins, max(1, episodes))
    reveal_den = float(max(1, reveal_total))
    return {                                                              # eval.py:490-510, key for key
        "win_rate": wins / max(1, episodes), "win_ci_low": ci_low, "win_ci_high": ci_high,
        "avg_steps": total_steps / max(1, episodes), "avg_progress": total_progress / max(1, episodes),
        "invalid_rate": invalids / max(1, total_steps),
        "forced_guess_rate": forced_guess_total / reveal_den,
        "forced_guess_success_rate": _ratio(forced_guess_success, forced_guess_total),
        "forced_guess_episode_rate": forced_guess_episodes / float(max(1, episodes)),
        "safe_option_rate": safe_option_total / reveal_den,
        "safe_option_miss_rate": _ratio(safe_option_misses, safe_option_total),
        "safe_option_pick_rate": _ratio(safe_option_hits, safe_option_total),
        "avg_safe_options_per_turn": _ratio(safe_cells, safe_option_total),
        "avg_frontier_component_size": _ratio(float(sum(component_sizes)), len(component_sizes)),
        "avg_selected_component_size": _ratio(float(sum(chosen_sizes)), len(chosen_sizes)),
        "belief_auroc": _auroc(y, p) if len(p) else float("nan"),
        "belief_ece": _ece(p, y) if len(p) else float("nan"),
        "wins": float(wins), "episodes": float(episodes),
        # extra (not in the reference dict): the subset-rule step statistics eval.py:362-379 accumulates
        "forced_step_rate": forced_steps / max(1, forced_steps + guess_attempts),
        "forced_step_accuracy": forced_correct / max(1, forced_steps),
        "guess_success_rate": guess_success / max(1, guess_attempts),
    }
