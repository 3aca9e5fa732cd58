"""Parity of the CUDA env (through the C ABI) with (a) fixtures recorded from the live reference
and (b) the CPU oracle at BASELINE.json sizes.  Every comparison is bit-exact.  Needs a B200."""
from types import SimpleNamespace as NS

import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback to test)")
    return torch


@pytest.mark.parametrize("name", P.ENV_CASES)
def test_cuda_env_matches_reference_fixture(torch_cuda, name):
    P.replay_env_case(name, lambda cfg, N: P.CudaAdapter(cfg, N))


def test_numpy_api_is_reference_shaped(torch_cuda):
    """api='numpy' (the reference calling convention, msw_step_host underneath): NumPy in/out and
    list-of-dict infos identical to env.py:485-505."""
    import minesweeper_ppo_b200 as m
    g = P.load("env_16x16x40_any")
    cfg = P.cfg_of(g)
    N, T, HW = int(g["N"]), int(g["T"]), cfg.H * cfg.W
    v = m.VecMinesweeper(N, m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4), seed=0)
    b = v.reset()
    assert isinstance(b["obs"], np.ndarray) and b["obs"].dtype == np.float32 and b["action_mask"].dtype == bool
    assert v.action_space() == 256 and v.obs_channels() == 10 and len(v.envs) == N
    names = (None, "win", "loss")
    for t in range(T):
        v.inject_layouts(*P.injections(g, t, N, HW))
        batch, rew, done, infos = v.step(g["actions"][t])
        P.assert_bits_equal(batch["obs"].reshape(N, -1), P.unpack(g["obs"][t], 10 * HW).astype(np.float32), f"t={t} obs")
        P.assert_bits_equal(batch["action_mask"], P.unpack(g["mask"][t], HW), f"t={t} mask")
        P.assert_bits_equal(rew, g["rewards"][t], f"t={t} rewards")
        P.assert_bits_equal(done, g["dones"][t], f"t={t} dones")
        assert infos["outcome"] == [names[o] for o in g["outcome"][t]]
        assert infos["done"] == [bool(d) for d in g["dones"][t]]
        for i in (0, N // 2, N - 1):
            assert infos["aux"][i] == {"step": int(g["step"][t][i]), "last_new_reveals": int(g["new_reveals"][t][i]),
                                       "revealed_frac": float(g["revealed_frac"][t][i])}
        if t % 16 == 0:          # the vec.envs[i] views eval.py:350-398 reads
            e = v.envs[3]
            P.assert_bits_equal(e.revealed.reshape(-1), P.unpack(g["st_revealed"][t][3], HW), "envs[3].revealed")
            P.assert_bits_equal(e.mine_mask.reshape(-1), P.unpack(g["st_mine"][t][3], HW), "envs[3].mine_mask")
            P.assert_bits_equal(e.adjacent_counts.reshape(-1), g["st_counts"][t][3], "envs[3].adjacent_counts")
            assert e.first_click_done == bool(g["st_first"][t][3]) and e.step_count == int(g["st_step_count"][t][3])
            assert not e.flags.any() and e.H == 16 and e.W == 16


@pytest.mark.parametrize("H,W,M", [(16, 16, 40), (16, 30, 99), (9, 9, 10)])
def test_numpy_api_recycled_result_arrays(torch_cuda, oracle, H, W, M):
    """The NumPy convention recycles its result arrays and rewrites only what changed since a set was last filled
    (msw_host_out.shadow).  Whatever the caller does with earlier results -- drops them at once (one set reused
    every step), keeps the previous batch (two sets alternate), keeps views alive for a while (more sets, stale by
    several steps), or passes its own `out` arrays -- every result equals the oracle's, and results the caller
    still holds are never written to."""
    import os
    import minesweeper_ppo_b200 as m
    N, HW = 333, H * W
    cfg = m.EnvConfig(H=H, W=W, mine_count=M, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    v = m.VecMinesweeper(N, cfg, seed=11)
    cpu = oracle.OracleVecEnv(N, cfg, seed=11, nthreads=os.cpu_count() or 1)
    rng = np.random.default_rng(3)
    batch = v.reset()
    b = cpu.reset()
    P.assert_bits_equal(batch["obs"], b["obs"], "reset obs")
    held = []                                                   # (release step, view, snapshot)
    for t in range(48):
        sc = rng.random((N, HW))
        sc[~b["action_mask"]] = -1
        a = sc.argmax(1) if t % 6 != 5 else rng.integers(0, HW, size=N)
        phase = t // 12
        v.host_delta = not (phase == 1 and t % 4 == 1)          # a full rewrite now and then: its set's shadow goes stale
        if phase == 0:
            batch = None                                        # nothing referenced: one set, one step stale
        if phase == 2 and t % 3 == 0:
            held.append((t + 4, batch["obs"][5:9], batch["obs"][5:9].copy()))      # a VIEW keeps the set out of the pool
        if phase == 3 and t % 2 == 0:
            own = (np.full((N, 10, H, W), np.nan, np.float32), np.zeros((N, HW), bool))
            pin = v._host_buffers()[0]
            pin["actions"].numpy()[:] = a
            res = v.step_host(pin["actions"], out=own)
            batch = {"obs": res["obs"], "action_mask": res["mask"]}
            assert batch["obs"] is own[0]
        else:
            batch, _, _, _ = v.step(a)
        b, _, _, _ = cpu.step(a, tensor_infos=True)
        P.assert_bits_equal(batch["obs"], b["obs"], f"t={t} obs")
        P.assert_bits_equal(batch["action_mask"], b["action_mask"], f"t={t} mask")
        for rel, view, snap in held:
            assert np.array_equal(view, snap), f"a held result was overwritten at t={t}"
        held = [h for h in held if h[0] > t]
    assert len(v._result_pool) >= 2 and any(e.valid for e in v._result_pool)
    assert int(cpu.episode_idx.max()) > 1


def test_numpy_api_sliced_state_copy_at_scale(torch_cuda, oracle):
    """From 32,768 envs on msw_step_host copies the packed state in four slices and expands slice k while the later
    ones are still on the bus; 40,000 envs (ragged last slice) against the oracle, delta and full-rewrite modes."""
    import os
    import minesweeper_ppo_b200 as m
    N, H, W, HW = 40000, 16, 16, 256
    cfg = m.EnvConfig(H=H, W=W, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    v = m.VecMinesweeper(N, cfg, seed=21)
    cpu = oracle.OracleVecEnv(N, cfg, seed=21, nthreads=os.cpu_count() or 1)
    rng = np.random.default_rng(4)
    batch, b = v.reset(), cpu.reset()
    for t in range(7):
        sc = rng.random((N, HW), dtype=np.float32)
        sc[~b["action_mask"]] = -1
        a = sc.argmax(1)
        v.host_delta = t != 4
        batch, r, d, infos = v.step(a)
        b, r0, d0, _ = cpu.step(a, tensor_infos=True)
        P.assert_bits_equal(batch["obs"], b["obs"], f"t={t} obs")
        P.assert_bits_equal(batch["action_mask"], b["action_mask"], f"t={t} mask")
        P.assert_bits_equal(r, r0, f"t={t} rewards")
        P.assert_bits_equal(d, d0, f"t={t} dones")
        assert len(infos["aux"]) == N and infos["done"][N - 1] == bool(d0[N - 1])


def test_cuda_floodfill_with_flags_matches_numba(torch_cuda):
    """Arbitrary (mines, revealed, flags, start) boards, including flags the hot path never sets."""
    import minesweeper_ppo_b200 as m
    g = P.load("floodfill")
    for H, W in g["shapes"]:
        H, W = int(H), int(W)
        HW = H * W
        inp = P.unpack(g[f"in_{H}x{W}"], 3 * HW)
        want = P.unpack(g[f"rev_{H}x{W}"], HW)
        rcn = g[f"rcn_{H}x{W}"]
        n = inp.shape[0]
        mines, rev, flags = (inp[:, i * HW:(i + 1) * HW] for i in range(3))
        v = m.VecMinesweeper(n, m.EnvConfig(H=H, W=W, mine_count=0), api="torch")
        v.reset()
        v.set_state(mine=mines, revealed=rev, flags=flags, first_click_done=np.ones(n, np.int32))
        cells = rcn[:, 0] * W + rcn[:, 1]
        _, _, dones, info = v.step(torch_cuda.from_numpy(cells.astype(np.int64)))
        dones = dones.cpu().numpy()
        new = info["last_new_reveals"].cpu().numpy()
        after = v._unpacked()["revealed"].astype(bool)
        start_mine = mines[np.arange(n), cells]
        start_rev = rev[np.arange(n), cells]
        for k in range(n):
            if start_mine[k] and not start_rev[k]:
                assert dones[k]                    # step() reports the loss; numba is never called here
                continue
            assert new[k] == rcn[k, 2], (H, W, k)
            if not dones[k]:
                P.assert_bits_equal(after[k], want[k], f"floodfill {H}x{W} #{k}")


SCALE_CFGS = [
    ("16x16x40", NS(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, win_reward=1.0,
                    loss_reward=-1.0, step_penalty=1e-4), 65536, 24),
    ("16x30x99", NS(H=16, W=30, mine_count=99, guarantee_safe_neighborhood=True, win_reward=1.0,
                    loss_reward=-1.0, step_penalty=1e-4), 16384, 16),
    ("9x9x10", NS(H=9, W=9, mine_count=10, guarantee_safe_neighborhood=True, win_reward=1.0,
                  loss_reward=-1.0, step_penalty=1e-4), 4096, 24),
    ("8x8x50dense", NS(H=8, W=8, mine_count=50, guarantee_safe_neighborhood=True, win_reward=1.0,
                       loss_reward=-1.0, step_penalty=1e-4), 4096, 12),
    ("32x32x200", NS(H=32, W=32, mine_count=200, guarantee_safe_neighborhood=False, win_reward=1.0,
                     loss_reward=-1.0, step_penalty=1e-4), 2048, 12),
]


@pytest.mark.parametrize("name,cfg,N,T", SCALE_CFGS, ids=[c[0] for c in SCALE_CFGS])
def test_cuda_env_matches_oracle_at_scale(torch_cuda, oracle, name, cfg, N, T):
    """BASELINE.json config sizes (65,536 envs at 16x16x40): CUDA and oracle both DRAW their own
    boards with the shared counter-based sampler spec, so this also pins the device sampler."""
    import os
    torch = torch_cuda
    seed, base = 1234, 7_000_000_000          # env ids beyond 2^32 exercise the 64-bit counter
    gpu = P.CudaAdapter(cfg, N, seed=seed, env_id_base=base)
    cpu = oracle.OracleVecEnv(N, cfg, seed=seed, env_id_base=base, nthreads=os.cpu_count() or 1, aux_maps=True)
    o_g, m_g = gpu.reset()
    b = cpu.reset()
    P.assert_bits_equal(o_g, b["obs"], "reset obs")
    P.assert_bits_equal(m_g, b["action_mask"], "reset mask")
    HW = cfg.H * cfg.W
    for t in range(T):
        valid_only = (t % 3) != 2
        a = gpu.v.random_actions(step_index=t, valid_only=valid_only, seed=99)
        a_np = a.cpu().numpy().astype(np.int64)
        assert a_np.min() >= 0 and a_np.max() < HW
        if valid_only:                        # synthetic source must pick unrevealed cells
            assert cpu.revealed[np.arange(N), a_np].sum() == 0
        g = gpu.step(a)
        bc, r, d, info = cpu.step(a_np, tensor_infos=True)
        tag = f"{name} t={t}"
        P.assert_bits_equal(g["rewards"], r, tag + " rewards")
        P.assert_bits_equal(g["dones"], d, tag + " dones")
        P.assert_bits_equal(g["outcome"], info["outcome_code"], tag + " outcome")
        P.assert_bits_equal(g["new_reveals"], info["last_new_reveals"], tag + " new_reveals")
        P.assert_bits_equal(g["step"], info["step"], tag + " step")
        P.assert_bits_equal(g["revealed_count"], info["revealed_count"], tag + " revealed_count")
        P.assert_bits_equal(g["mask"], bc["action_mask"], tag + " mask")
        P.assert_bits_equal(g["obs"], bc["obs"], tag + " obs")
        P.assert_bits_equal(g["labels"], cpu.mine_labels, tag + " labels")
        P.assert_bits_equal(g["valid"], cpu.mine_valid, tag + " valid")
        if t % 4 == 3 or t == T - 1:
            s = gpu.state()
            P.assert_bits_equal(s["mine"], cpu.mine.astype(bool), tag + " mines")
            P.assert_bits_equal(s["revealed"], cpu.revealed.astype(bool), tag + " revealed")
            P.assert_bits_equal(s["counts"], cpu.counts, tag + " counts")
            P.assert_bits_equal(s["first"], cpu.first_click_done.astype(bool), tag + " first")
    assert int(cpu.episode_idx.max()) > 1     # auto-resets happened


def test_size_independent_properties_full_size(torch_cuda):
    """C2 size on device only: properties every observation must satisfy (env.py:172-196)."""
    import minesweeper_ppo_b200 as m
    torch = torch_cuda
    N = 65536
    v = m.VecMinesweeper(N, m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4), seed=5, api="torch", aux_maps=True)
    b = v.reset()
    assert float(b["obs"].abs().sum()) == 0.0 and bool(b["action_mask"].all())
    for t in range(40):
        a = v.random_actions(t)
        b, r, d, info = v.step(a)
        obs, mask = b["obs"], b["action_mask"]
        assert bool(((obs == 0) | (obs == 1)).all())
        rev = obs[:, 0].reshape(N, -1)
        assert bool((mask == (rev == 0)).all())                          # mask = ~revealed
        assert bool((obs[:, 1:].sum(1).reshape(N, -1) == rev).all())     # exactly one count plane per revealed cell
        st = v.state_tensors
        pop = sum(((st["revealed"] >> k) & 1).sum(1) for k in range(32))
        assert bool((pop == rev.sum(1).to(pop.dtype)).all())            # obs ch0 == bitboard popcount
        first = st["meta"][:, 0] == 1
        mines = sum(((st["mines"] >> k) & 1).sum(1) for k in range(32))
        assert bool((mines[first] == 40).all()) and bool((mines[~first] == 0).all())
        assert bool((v.mine_labels.reshape(N, -1).sum(1) == mines).all())
        assert bool((v.mine_valid.reshape(N, -1)[first] == mask[first]).all())
        assert not bool(v.mine_valid.reshape(N, -1)[~first].any())
        # loss <=> reward == loss constant; no revealed mine survives the auto-reset
        assert bool(((r == float(v.reward_constants[1])) == (info["outcome_code"] == 2)).all())
        assert int((st["revealed"] & st["mines"]).ne(0).sum()) == 0


def test_sampler_uniformity_chi2(torch_cuda):
    """Device board sampler: exactly mine_count mines, none in the 3x3 around the first click,
    per-cell marginals uniform (chi-square) -- SURVEY 4.2."""
    import minesweeper_ppo_b200 as m
    from scipy import stats
    torch = torch_cuda
    N = 262144
    v = m.VecMinesweeper(N, m.EnvConfig(H=16, W=16, mine_count=40), seed=77, api="torch")
    v.reset()
    click = 7 * 16 + 8
    v.step(torch.full((N,), click, dtype=torch.int32, device=v.device))
    u = v._unpacked()["mine"].astype(np.int64)
    assert (u.sum(1) == 40).all()
    grid = u.sum(0).reshape(16, 16)
    assert grid[6:9, 7:10].sum() == 0
    allowed = np.ones((16, 16), bool)
    allowed[6:9, 7:10] = False
    obs_counts = grid[allowed]
    expected = N * 40 / allowed.sum()
    chi2 = ((obs_counts - expected) ** 2 / (expected * (1 - 40 / allowed.sum()))).sum()
    p = stats.chi2.sf(chi2, df=allowed.sum() - 1)
    assert p > 1e-4, (chi2, p)


def test_shard_invariance(torch_cuda):
    """Env i behaves the same whether it lives in one shard or in the second of two."""
    import minesweeper_ppo_b200 as m
    torch = torch_cuda
    cfg = m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4)
    N = 2048
    whole = m.VecMinesweeper(N, cfg, seed=3, api="torch")
    lo = m.VecMinesweeper(N // 2, cfg, seed=3, api="torch", env_id_base=0)
    hi = m.VecMinesweeper(N // 2, cfg, seed=3, api="torch", env_id_base=N // 2)
    for v in (whole, lo, hi):
        v.reset()
    for t in range(24):
        a = whole.random_actions(t)
        b, r, d, _ = whole.step(a)
        b0, r0, d0, _ = lo.step(a[: N // 2].contiguous())
        b1, r1, d1, _ = hi.step(a[N // 2:].contiguous())
        assert torch.equal(b["obs"], torch.cat([b0["obs"], b1["obs"]]))
        assert torch.equal(r, torch.cat([r0, r1])) and torch.equal(d, torch.cat([d0, d1]))


def test_out_buffers_and_bad_arguments(torch_cuda):
    import minesweeper_ppo_b200 as m
    torch = torch_cuda
    cfg = m.EnvConfig(H=16, W=16, mine_count=40)
    v = m.VecMinesweeper(64, cfg, seed=1, api="torch")
    buf = m.RolloutBuffer(64, 4, (10, 16, 16), 256, v.device, aux_maps=True)
    v.aux_maps = True
    v.reset(out=buf.slot(0))
    assert bool(buf.action_mask[:64].all()) and float(buf.obs[:64].sum()) == 0
    a = v.random_actions(0)
    nxt, cur = buf.slot(1), buf.slot(0)
    out = m.StepOut(obs=nxt.obs, action_mask=nxt.action_mask, rewards=cur.rewards, dones=cur.dones,
                    mine_labels=nxt.mine_labels, mine_valid=nxt.mine_valid)
    b, r, d, _ = v.step(a, out=out)
    assert b["obs"].data_ptr() == buf.obs[64:128].data_ptr() and r.data_ptr() == buf.rewards.data_ptr()
    assert float(buf.obs[64:128].sum()) > 0 and float(buf.rewards[:64].abs().sum()) > 0
    with pytest.raises((ValueError, AssertionError)):
        v.step(torch.zeros(63, dtype=torch.int32, device=v.device))
    with pytest.raises(ValueError):
        v.step(a, out=m.StepOut(obs=torch.empty(64, 10, 16, 16, device=v.device, dtype=torch.float16),
                                action_mask=nxt.action_mask))
    with pytest.raises(ValueError):
        m.VecMinesweeper(4, m.EnvConfig(H=40, W=40, mine_count=10))


def test_builtin_synthetic_policy_equals_separate_action_source(torch_cuda):
    """msw_step with rand_mode draws exactly the action msw_random_actions would have produced."""
    import minesweeper_ppo_b200 as m
    torch = torch_cuda
    for H, W, M in [(16, 16, 40), (16, 30, 99), (9, 9, 10)]:
        cfg = m.EnvConfig(H=H, W=W, mine_count=M, step_penalty=1e-4)
        a_env = m.VecMinesweeper(3000, cfg, seed=5, api="torch", aux_maps=True)
        b_env = m.VecMinesweeper(3000, cfg, seed=5, api="torch", aux_maps=True)
        a_env.reset(); b_env.reset()
        acts = torch.empty(3000, dtype=torch.int32, device=a_env.device)
        for t in range(20):
            valid_only = t % 4 != 3
            want = a_env.random_actions(t, valid_only=valid_only, seed=42)
            ba, ra, da, _ = a_env.step(want)
            bb, rb, db, _ = b_env.step_random(t, valid_only=valid_only, seed=42, actions_out=acts)
            assert torch.equal(acts, want), (H, W, t)
            assert torch.equal(ba["obs"], bb["obs"]) and torch.equal(ba["action_mask"], bb["action_mask"])
            assert torch.equal(ra, rb) and torch.equal(da, db)
            assert torch.equal(a_env.mine_labels, b_env.mine_labels)


def test_random_board_shapes_against_oracle(torch_cuda, oracle):
    """Sweep of random (H, W, mines, safe-neighbourhood) combinations -- including W in {1, 31, 32},
    planes that are not 16-byte multiples (scalar encoder) and dense boards (complement sampling) --
    each stepped against the oracle, everything compared bit for bit."""
    import os
    rng = np.random.default_rng(2024)
    shapes = [(1, 12), (12, 1), (2, 2), (3, 31), (31, 3), (32, 32), (7, 32), (32, 7), (5, 5), (10, 10), (24, 30)]
    while len(shapes) < 26:
        W = int(rng.integers(1, 33)); H = int(rng.integers(1, 1024 // W + 1))
        shapes.append((min(H, 40), W))
    for k, (H, W) in enumerate(shapes):
        HW = H * W
        if HW < 2:
            continue
        M = int(rng.integers(0, HW))                 # 0 .. HW-1, so sparse, dense and mine-free boards all occur
        cfg = NS(H=H, W=W, mine_count=M, guarantee_safe_neighborhood=bool(k % 3), win_reward=1.0,
                 loss_reward=-1.0, step_penalty=1e-4)
        N = 257
        gpu = P.CudaAdapter(cfg, N, seed=k, env_id_base=k * 1000)
        cpu = oracle.OracleVecEnv(N, cfg, seed=k, env_id_base=k * 1000, aux_maps=True, nthreads=os.cpu_count() or 1)
        o, mk = gpu.reset(); b = cpu.reset()
        P.assert_bits_equal(o, b["obs"], f"{H}x{W}x{M} reset obs")
        for t in range(8):
            a = gpu.v.random_actions(t, valid_only=bool(t % 2), seed=k)
            g = gpu.step(a)
            bc, r, d, info = cpu.step(a.cpu().numpy().astype(np.int64), tensor_infos=True)
            tag = f"{H}x{W}x{M} t={t}"
            P.assert_bits_equal(g["obs"], bc["obs"], tag + " obs")
            P.assert_bits_equal(g["mask"], bc["action_mask"], tag + " mask")
            P.assert_bits_equal(g["rewards"], r, tag + " rewards")
            P.assert_bits_equal(g["dones"], d, tag + " dones")
            P.assert_bits_equal(g["outcome"], info["outcome_code"], tag + " outcome")
            P.assert_bits_equal(g["new_reveals"], info["last_new_reveals"], tag + " new")
            P.assert_bits_equal(g["labels"], cpu.mine_labels, tag + " labels")
            P.assert_bits_equal(g["valid"], cpu.mine_valid, tag + " valid")
        s = gpu.state()
        P.assert_bits_equal(s["mine"], cpu.mine.astype(bool), f"{H}x{W}x{M} mines")
        P.assert_bits_equal(s["counts"], cpu.counts, f"{H}x{W}x{M} counts")
