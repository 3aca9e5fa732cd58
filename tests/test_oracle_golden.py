"""Pins the CPU oracle (oracle/msw_oracle.c) against fixtures recorded from the live
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import parity as P


@pytest.mark.parametrize("name", P.ENV_CASES)
def test_oracle_env_matches_reference(oracle, name):
    P.replay_env_case(name, lambda cfg, N: P.OracleAdapter(oracle, cfg, N))


def test_oracle_env_threads_equal_serial(oracle):
    P.replay_env_case("16x16x40_valid", lambda cfg, N: P.OracleAdapter(oracle, cfg, N, nthreads=4))


def test_oracle_floodfill_matches_numba_and_python(oracle):
    g = P.load("floodfill")
    for H, W in g["shapes"]:
        HW = int(H) * int(W)
        inp = P.unpack(g[f"in_{H}x{W}"], 3 * HW)
        want = P.unpack(g[f"rev_{H}x{W}"], HW)
        rcn = g[f"rcn_{H}x{W}"]
        for k in range(inp.shape[0]):
            mines, rev, flags = (inp[k, i * HW:(i + 1) * HW].reshape(H, W) for i in range(3))
            counts = oracle.adjacent_counts(mines)
            got, n = oracle.flood_fill(rev, flags, mines, counts, rcn[k, 0], rcn[k, 1])
            assert n == rcn[k, 2], (H, W, k)
            P.assert_bits_equal(got.reshape(-1), want[k], f"floodfill {H}x{W} #{k}")


def test_oracle_gae_bit_exact(oracle):
    g = P.load("gae")
    for name in g["names"]:
        gamma, lam = g[f"{name}_gamma_lam"]
        if bool(g[f"{name}_lv_is_fp16"]):
            continue          # fp16 bootstrap pre-rounding is a wrapper concern (see test_gae)
        adv, ret = oracle.gae(g[f"{name}_rewards"], g[f"{name}_values"], g[f"{name}_dones"],
                              g[f"{name}_last_values"], float(gamma), float(lam))
        P.assert_bits_equal(adv, g[f"{name}_adv"], f"gae {name} adv")
        P.assert_bits_equal(ret, g[f"{name}_ret"], f"gae {name} ret")


def test_oracle_rollout_buffer_protocol(oracle):
    """Replays the real collect_rollout (train_rl.py:155-289): slot t holds obs_t/mask_t/
    labels_t/valid_t (state BEFORE a_t) and reward_t/done_t (result OF a_t)."""
    g = P.load("rollout_16x16x40")
    cfg = P.cfg_of(g)
    N, T, HW = int(g["N"]), int(g["T"]), cfg.H * cfg.W
    env = P.OracleAdapter(oracle, cfg, N)
    obs, mask = env.reset()
    lab, val = env.v.mine_labels, env.v.mine_valid
    for t in range(T):
        P.assert_bits_equal(obs.reshape(N, -1), P.unpack(g["obs"][t], 10 * HW).astype(np.float32), f"slot {t} obs")
        P.assert_bits_equal(mask, P.unpack(g["mask"][t], HW), f"slot {t} mask")
        P.assert_bits_equal(lab.reshape(N, -1), P.unpack(g["mine_labels"][t], HW).astype(np.float32), f"slot {t} labels")
        P.assert_bits_equal(val.reshape(N, -1), P.unpack(g["mine_valid"][t], HW), f"slot {t} valid")
        mine, sel = P.injections(g, t, N, HW)
        o = env.step(g["actions"][t], mine, sel)
        P.assert_bits_equal(o["rewards"], g["rewards"][t], f"slot {t} rewards")
        P.assert_bits_equal(o["dones"], g["dones"][t], f"slot {t} dones")
        obs, mask, lab, val = o["obs"], o["mask"], o["labels"], o["valid"]
    adv, ret = oracle.gae(g["rewards"], g["values"], g["dones"], g["last_values"], 0.995, 0.95)
    P.assert_bits_equal(adv, g["adv"], "rollout adv")
    P.assert_bits_equal(ret, g["ret"], "rollout ret")


def test_oracle_sampler_properties(oracle):
    """Board sampler (not in the parity contract): exact count, safe zone, fallback."""
    from types import SimpleNamespace as NS
    cfg = NS(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, win_reward=1.0,
             loss_reward=-1.0, step_penalty=1e-4)
    seen = set()
    for e in range(200):
        r0, c0 = e % 16, (e * 7) % 16
        m = oracle.place_mines(cfg, seed=0, env_id=e, episode=1, r0=r0, c0=c0)
        assert m.sum() == 40
        assert not m[max(0, r0 - 1):r0 + 2, max(0, c0 - 1):c0 + 2].any()
        seen.add(m.tobytes())
    assert len(seen) == 200
    small = NS(H=4, W=4, mine_count=8, guarantee_safe_neighborhood=True, win_reward=1.0,
               loss_reward=-1.0, step_penalty=1e-4)
    m = oracle.place_mines(small, 0, 0, 0, 1, 1)          # 7 allowed < 8 -> only (1,1) is safe
    assert m.sum() == 8 and not m[1, 1]
    dense = NS(H=8, W=8, mine_count=50, guarantee_safe_neighborhood=True, win_reward=1.0,
               loss_reward=-1.0, step_penalty=1e-4)
    m = oracle.place_mines(dense, 3, 9, 2, 4, 4)          # complement mode
    assert m.sum() == 50 and not m[3:6, 3:6].any()
