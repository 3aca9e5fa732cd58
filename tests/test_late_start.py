"""Late-start curriculum (env.py:416-466).

The reference draws from ONE sequential NumPy generator shared by all envs, so bit parity with it
is impossible for any parallel implementation; it is pinned in two steps instead:
  * CPU: the oracle's statistics match statistics recorded from the live reference
    (tests/golden/late_start_stats.json, written by make_golden.py) within sampling error;
  * GPU: the CUDA kernel equals the oracle bit for bit (same counter-based stream)."""
import json
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest

import parity as P

STATS = json.load(open(os.path.join(P.GOLDEN, "late_start_stats.json")))


def _tuple(ls, HW, seed=12):
    lo = max(1, int(ls.get("min_hidden", 5)))
    hi = max(lo, int(ls.get("max_hidden", lo)))
    return (seed, float(ls["prob"]), lo, hi, max(1, int(ls.get("max_attempts", 3))),
            max(1, int(ls.get("max_extra_steps", HW))))


def _cfg(c):
    return NS(H=c["H"], W=c["W"], mine_count=c["mine_count"], guarantee_safe_neighborhood=True,
              win_reward=1.0, loss_reward=-1.0, step_penalty=1e-4)


def _stats(v, cfg):
    safe = cfg.H * cfg.W - cfg.mine_count
    f = v.first_click_done.astype(bool)
    hid = safe - (v.revealed.astype(bool) & ~v.mine.astype(bool)).sum(1)
    return dict(frac_started=float(f.mean()), mean_safe_hidden_started=float(hid[f].mean()),
                max_safe_hidden_started=int(hid[f].max()), mean_step_count_started=float(v.step_count[f].mean()))


@pytest.mark.parametrize("name", sorted(STATS))
def test_oracle_late_start_statistics_match_reference(oracle, name):
    ref = STATS[name]
    cfg = _cfg(ref["cfg"])
    HW = cfg.H * cfg.W
    N = 8192
    v = oracle.OracleVecEnv(N, cfg, seed=5, late_start=_tuple(ref["late_start"], HW), nthreads=os.cpu_count() or 1)
    mask = v.reset()["action_mask"]
    s0 = _stats(v, cfg)
    r0 = ref["after_reset"]
    assert not (v.revealed.astype(bool) & v.mine.astype(bool)).any()          # only safe cells are pre-played
    assert s0["max_safe_hidden_started"] <= ref["late_start"].get("max_hidden", ref["late_start"].get("min_hidden", 5))
    assert abs(s0["frac_started"] - r0["frac_started"]) < 0.03
    assert abs(s0["mean_safe_hidden_started"] - r0["mean_safe_hidden_started"]) < 0.05 * r0["mean_safe_hidden_started"] + 0.1
    assert abs(s0["mean_step_count_started"] - r0["mean_step_count_started"]) < 0.05 * r0["mean_step_count_started"] + 0.1
    rng = np.random.default_rng(1)
    wins = losses = 0
    for t in range(40):                      # auto-resets re-apply the late start (env.py:497-498 -> :406-414)
        s = rng.random(mask.shape); s[~mask] = -1
        b, r, d, info = v.step(s.argmax(1), tensor_infos=True)
        mask = b["action_mask"]
        wins += int((info["outcome_code"] == 1).sum()); losses += int((info["outcome_code"] == 2).sum())
    s1, r1 = _stats(v, cfg), ref["after_40_steps"]
    assert abs(s1["frac_started"] - r1["frac_started"]) < 0.04
    assert abs(s1["mean_safe_hidden_started"] - r1["mean_safe_hidden_started"]) < 0.08 * r1["mean_safe_hidden_started"] + 0.2
    assert abs(wins / max(1, wins + losses) - ref["win_rate_random_play"]) < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(STATS) + ["16x30x99"])
def test_cuda_late_start_equals_oracle(oracle, name):
    import torch
    import minesweeper_ppo_b200 as m
    if name == "16x30x99":
        c, ls = dict(H=16, W=30, mine_count=99), dict(prob=0.5, min_hidden=10, max_hidden=60)
    else:
        c, ls = STATS[name]["cfg"], STATS[name]["late_start"]
    cfg = _cfg(c)
    HW, N, seed, base = cfg.H * cfg.W, 4096, 21, 5_000_000_000
    gpu = m.VecMinesweeper(N, m.EnvConfig(H=cfg.H, W=cfg.W, mine_count=cfg.mine_count, step_penalty=1e-4), seed=seed,
                           late_start_cfg=ls, late_start_seed=77, api="torch", aux_maps=True, env_id_base=base)
    cpu = oracle.OracleVecEnv(N, cfg, seed=seed, env_id_base=base, late_start=_tuple(ls, HW, seed=77), aux_maps=True,
                              nthreads=os.cpu_count() or 1)
    bg, bc = gpu.reset(), cpu.reset()
    P.assert_bits_equal(bg["obs"].cpu().numpy(), bc["obs"], "reset obs")
    P.assert_bits_equal(bg["action_mask"].cpu().numpy(), bc["action_mask"], "reset mask")
    assert float(bg["obs"].sum()) > 0
    for t in range(30):
        a = gpu.random_actions(t, valid_only=(t % 5 != 4), seed=9)
        bg, rg, dg, ig = gpu.step(a)
        bc, rc, dc, ic = cpu.step(a.cpu().numpy(), tensor_infos=True)
        tag = f"{name} t={t}"
        P.assert_bits_equal(rg.cpu().numpy(), rc, tag + " rewards")
        P.assert_bits_equal(dg.cpu().numpy(), dc, tag + " dones")
        P.assert_bits_equal(ig["step"].cpu().numpy(), ic["step"], tag + " aux.step")
        P.assert_bits_equal(bg["obs"].cpu().numpy(), bc["obs"], tag + " obs")
        P.assert_bits_equal(bg["action_mask"].cpu().numpy(), bc["action_mask"], tag + " mask")
        P.assert_bits_equal(gpu.mine_labels.cpu().numpy(), cpu.mine_labels, tag + " labels")
    u = gpu._unpacked()
    P.assert_bits_equal(u["mine"], cpu.mine, "mines")
    P.assert_bits_equal(u["revealed"], cpu.revealed, "revealed")
    P.assert_bits_equal(u["meta"][:, 1], cpu.step_count, "step_count")
    P.assert_bits_equal(u["meta"][:, 2].astype(np.uint32), cpu.episode_idx, "episode_idx")


@pytest.mark.gpu
def test_numpy_api_with_late_start(oracle):
    import minesweeper_ppo_b200 as m
    ls = dict(prob=1.0, min_hidden=2, max_hidden=6)
    v = m.VecMinesweeper(64, m.EnvConfig(), seed=1, late_start_cfg=ls, late_start_seed=3)
    o = oracle.OracleVecEnv(64, oracle.OracleEnvConfig(), seed=1, late_start=_tuple(ls, 64, seed=3))
    b, c = v.reset(), o.reset()
    P.assert_bits_equal(b["obs"], c["obs"], "reset obs")
    rng = np.random.default_rng(0)
    for t in range(10):
        s = rng.random(b["action_mask"].shape); s[~b["action_mask"]] = -1
        a = s.argmax(1).astype(np.int32)
        b, r, d, info = v.step(a)
        c, r2, d2, info2 = o.step(a)
        P.assert_bits_equal(b["obs"], c["obs"], f"t={t} obs")
        P.assert_bits_equal(r, r2, "rewards")
        assert info["outcome"] == info2["outcome"] and info["aux"] == info2["aux"]
