#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the authoring container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

It imports the unmodified reference from /root/reference (read-only; numba's cache is
redirected to a temp dir), drives `VecMinesweeper` (minesweeper/env.py:379-517),
`flood_fill_reveal` (minesweeper/env_numba.py:16-77) and its pure-Python twin
(env.py:217-244), `RolloutBuffer.compute_gae` (minesweeper/buffers.py:78-94) and the
real `collect_rollout` (train_rl.py:155-289) on seeded inputs, and records inputs,
the mine layouts the reference drew (captured by wrapping -- not editing --
`MinesweeperEnv._place_mines_safe`, env.py:280-312) and every output.

Everything boolean / {0.0,1.0}-valued is stored bit-packed (np.packbits, little bit
order over the flattened array) after asserting that it really only holds 0/1.
"""
from __future__ import annotations

import os
import sys
import tempfile

REF = os.environ.get("MSW_REFERENCE", "/root/reference")
os.environ.setdefault("NUMBA_CACHE_DIR", tempfile.mkdtemp(prefix="numba_cache_"))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from minesweeper import env as ref_env  # noqa: E402
from minesweeper.buffers import RolloutBuffer  # noqa: E402
from minesweeper.env import EnvConfig, MinesweeperEnv, VecMinesweeper  # noqa: E402
from minesweeper.env_numba import HAS_ENV_NUMBA, flood_fill_reveal  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
assert HAS_ENV_NUMBA, "fixtures must come from the numba flood fill (env_numba.py)"


def pack(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    u = a.astype(np.uint8)
    assert np.array_equal(u.astype(a.dtype), a) and u.max(initial=0) <= 1, "not a 0/1 array"
    return np.packbits(u.reshape(a.shape[0], -1), axis=1, bitorder="little")


class LayoutRecorder:
    """Wraps MinesweeperEnv._place_mines_safe and logs every layout it produces."""

    def __init__(self):
        self.log = []          # (env object, mine_mask copy)
        self._orig = MinesweeperEnv._place_mines_safe

    def __enter__(self):
        rec, orig = self, self._orig

        def wrapped(env, first_click_rc):
            orig(env, first_click_rc)
            rec.log.append((env, env.mine_mask.copy()))

        MinesweeperEnv._place_mines_safe = wrapped
        return self

    def __exit__(self, *exc):
        MinesweeperEnv._place_mines_safe = self._orig

    def drain(self, index_of):
        out = [(index_of[id(e)], m) for e, m in self.log]
        self.log = []
        return out


OUTCOME = {None: 0, "win": 1, "loss": 2}


def run_env_case(name, cfg: EnvConfig, N: int, T: int, policy: str, seed: int):
    """policy: 'valid' | 'any' | 'win' (scripted: click every safe cell)."""
    rng = np.random.default_rng(seed + 1000)
    HW = cfg.H * cfg.W
    with LayoutRecorder() as rec:
        vec = VecMinesweeper(N, cfg, seed=seed)
        index_of = {id(e): i for i, e in enumerate(vec.envs)}
        batch = vec.reset()
        assert not rec.log
        obs0, mask0 = batch["obs"], batch["action_mask"]
        perm = np.stack([rng.permutation(HW) for _ in range(N)])   # per-env click order ('win')
        acts, place_t, place_i, place_bits = [], [], [], []
        rec_obs, rec_mask, rec_rew, rec_done, rec_outcome = [], [], [], [], []
        rec_new, rec_step, rec_frac = [], [], []
        st_rev, st_mine, st_cnt, st_first, st_stepc = [], [], [], [], []
        lab, val = [], []
        mask = mask0
        for t in range(T):
            if policy == "valid":
                s = rng.random(mask.shape)
                s[~mask] = -1.0
                a = s.argmax(1).astype(np.int64)
            elif policy == "any":
                a = rng.integers(-3 * HW, 3 * HW, size=N, dtype=np.int64)
            elif policy == "win":
                a = np.zeros(N, np.int64)
                for i, e in enumerate(vec.envs):
                    if not e.first_click_done:
                        a[i] = int(rng.integers(0, HW))
                    else:
                        safe_hidden = (~e.mine_mask & ~e.revealed).reshape(-1)
                        order = perm[i]
                        a[i] = int(order[np.argmax(safe_hidden[order])])
            else:
                raise ValueError(policy)
            batch, rew, done, infos = vec.step(a.astype(np.int32) if policy != "any" else a)
            for i, m in rec.drain(index_of):
                place_t.append(t); place_i.append(i); place_bits.append(m.reshape(-1))
            acts.append(a)
            rec_obs.append(pack(batch["obs"])); rec_mask.append(pack(batch["action_mask"]))
            rec_rew.append(rew.copy()); rec_done.append(done.copy())
            rec_outcome.append(np.array([OUTCOME[o] for o in infos["outcome"]], np.int8))
            rec_new.append(np.array([x["last_new_reveals"] for x in infos["aux"]], np.int32))
            rec_step.append(np.array([x["step"] for x in infos["aux"]], np.int32))
            rec_frac.append(np.array([x["revealed_frac"] for x in infos["aux"]], np.float64))
            assert infos["done"] == [bool(d) for d in done]
            st_rev.append(pack(np.stack([e.revealed for e in vec.envs])))
            st_mine.append(pack(np.stack([e.mine_mask for e in vec.envs])))
            st_cnt.append(np.stack([e.adjacent_counts for e in vec.envs]).reshape(N, -1).copy())
            st_first.append(np.array([e.first_click_done for e in vec.envs], bool))
            st_stepc.append(np.array([e.step_count for e in vec.envs], np.int32))
            # auxiliary maps exactly as train_rl.py:205-212 derives them from vec.envs
            L = np.zeros((N, cfg.H, cfg.W), np.float32)
            V = np.zeros((N, cfg.H, cfg.W), bool)
            for i, e in enumerate(vec.envs):
                if e.first_click_done:
                    np.copyto(L[i], e.mine_mask, casting="unsafe")
                    np.logical_and(~e.revealed, ~e.flags, out=V[i])
            lab.append(pack(L)); val.append(pack(V))
            mask = batch["action_mask"]
    place_bits = np.stack(place_bits) if place_bits else np.zeros((0, HW), bool)
    rew = np.stack(rec_rew)
    done = np.stack(rec_done)
    out = dict(
        H=cfg.H, W=cfg.W, mine_count=cfg.mine_count, safe=int(cfg.guarantee_safe_neighborhood),
        win_reward=cfg.win_reward, loss_reward=cfg.loss_reward, step_penalty=cfg.step_penalty,
        N=N, T=T, seed=seed,
        obs0=pack(obs0), mask0=pack(mask0),
        actions=np.stack(acts),
        place_t=np.array(place_t, np.int32), place_i=np.array(place_i, np.int32),
        place_bits=np.packbits(place_bits.astype(np.uint8), axis=1, bitorder="little"),
        obs=np.stack(rec_obs), mask=np.stack(rec_mask), rewards=rew, dones=done,
        outcome=np.stack(rec_outcome), new_reveals=np.stack(rec_new), step=np.stack(rec_step),
        revealed_frac=np.stack(rec_frac),
        st_revealed=np.stack(st_rev), st_mine=np.stack(st_mine), st_counts=np.stack(st_cnt),
        st_first=np.stack(st_first), st_step_count=np.stack(st_stepc),
        mine_labels=np.stack(lab), mine_valid=np.stack(val),
    )
    np.savez_compressed(os.path.join(OUT, f"env_{name}.npz"), **out)
    oc = np.stack(rec_outcome)
    print(f"env_{name}: N={N} T={T} episodes={int(done.sum())} wins={int((oc == 1).sum())} "
          f"losses={int((oc == 2).sum())} placements={len(place_t)} "
          f"distinct_rewards={sorted(set(rew.view(np.uint32).reshape(-1).tolist()))}")


def run_floodfill_case():
    """Arbitrary (mines, revealed, flags, start) boards -- including non-empty flags,
    which the hot path never produces -- through BOTH reference flood fills."""
    rng = np.random.default_rng(7)
    recs = []
    shapes = [(16, 16), (16, 30), (30, 16), (9, 9), (8, 8), (5, 7), (1, 12), (32, 32), (3, 3)]
    for H, W in shapes:
        inputs, outs_rev, outs_n = [], [], []
        for k in range(60):
            dens = rng.choice([0.0, 0.05, 0.12, 0.16, 0.3])
            mines = rng.random((H, W)) < dens
            revealed = rng.random((H, W)) < rng.choice([0.0, 0.1, 0.5])
            flags = rng.random((H, W)) < rng.choice([0.0, 0.0, 0.1, 0.3])
            r, c = int(rng.integers(0, H)), int(rng.integers(0, W))
            env = MinesweeperEnv(EnvConfig(H=H, W=W, mine_count=0))
            counts = env._compute_adjacent_counts(mines).copy()
            rev_nb = revealed.copy()
            n_nb = int(flood_fill_reveal(rev_nb, flags, mines, counts, r, c))
            # the pure-Python twin (env.py:217-244); it does not early-out on a mine start,
            # so compare only where the numba contract (start not a mine) is met by step().
            env.mine_mask[:] = mines; env.adjacent_counts[:] = counts
            env.revealed[:] = revealed; env.flags[:] = flags
            n_py = env._reveal_with_flood_fill_python(r, c)
            if not mines[r, c]:
                assert n_py == n_nb and np.array_equal(env.revealed, rev_nb)
            inputs.append(np.concatenate([mines.reshape(-1), revealed.reshape(-1), flags.reshape(-1)]))
            outs_rev.append(rev_nb.reshape(-1)); outs_n.append((r, c, n_nb))
        recs.append((H, W, np.stack(inputs), np.stack(outs_rev), np.array(outs_n, np.int32)))
    out = {"shapes": np.array([(h, w) for h, w, *_ in recs], np.int32)}
    for h, w, i, o, n in recs:
        out[f"in_{h}x{w}"] = np.packbits(i.astype(np.uint8), axis=1, bitorder="little")
        out[f"rev_{h}x{w}"] = np.packbits(o.astype(np.uint8), axis=1, bitorder="little")
        out[f"rcn_{h}x{w}"] = n
    np.savez_compressed(os.path.join(OUT, "floodfill.npz"), **out)
    print("floodfill:", {f"{h}x{w}": int(n[:, 2].sum()) for h, w, _, _, n in recs})


def run_gae_cases():
    out = {}
    names = []

    def case(name, T, N, p_done, gamma, lam, seed, lv_dtype=torch.float32, pattern=None):
        g = torch.Generator().manual_seed(seed)
        consts = torch.tensor([-1e-4, -1.0 - 1e-4, 1.0 - 1e-4], dtype=torch.float64).float()
        buf = RolloutBuffer(N, T, (1, 1, 1), 1, torch.device("cpu"))
        dones = torch.rand((T, N), generator=g) < p_done
        if pattern == "all":
            dones[:] = True
        elif pattern == "none":
            dones[:] = False
        elif pattern == "last":
            dones[:] = False; dones[-1] = True
        rewards = torch.where(dones, consts[torch.randint(1, 3, (T, N), generator=g)], consts[0])
        if pattern == "gauss":
            rewards = torch.randn((T, N), generator=g)
        values = 0.5 * torch.randn((T, N), generator=g)
        last_values = (0.5 * torch.randn((N,), generator=g)).to(lv_dtype)
        buf.rewards[:] = rewards.reshape(-1); buf.values[:] = values.reshape(-1)
        buf.dones[:] = dones.reshape(-1)
        buf.compute_gae(last_values, gamma=gamma, lam=lam)
        names.append(name)
        out[f"{name}_rewards"] = rewards.numpy(); out[f"{name}_values"] = values.numpy()
        out[f"{name}_dones"] = dones.numpy()
        out[f"{name}_last_values"] = last_values.float().numpy()
        out[f"{name}_lv_is_fp16"] = np.array(lv_dtype == torch.float16)
        out[f"{name}_gamma_lam"] = np.array([gamma, lam], np.float64)
        out[f"{name}_adv"] = buf.advantages.view(T, N).numpy().copy()
        out[f"{name}_ret"] = buf.returns.view(T, N).numpy().copy()

    case("c3like", 128, 96, 0.15, 0.995, 0.95, 0)
    case("short", 1, 33, 0.3, 0.995, 0.95, 1)
    case("t2", 2, 7, 0.5, 0.99, 0.9, 2)
    case("alldone", 16, 40, 0.0, 0.995, 0.95, 3, pattern="all")
    case("nodone", 64, 40, 0.0, 0.995, 0.95, 4, pattern="none")
    case("lastdone", 32, 40, 0.0, 0.995, 0.95, 5, pattern="last")
    case("gauss", 200, 65, 0.1, 0.9, 0.8, 6, pattern="gauss")
    case("lam1", 50, 31, 0.05, 1.0, 1.0, 7, pattern="gauss")
    case("lam0", 50, 31, 0.05, 0.97, 0.0, 8, pattern="gauss")
    case("fp16_last", 64, 48, 0.15, 0.995, 0.95, 9, lv_dtype=torch.float16)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **out)
    print("gae:", names)


def run_rollout_case():
    """The real collect_rollout (train_rl.py:155-289) with a stub policy: pins the
    buffer-write protocol (time alignment of obs/mask/labels vs reward/done) and the
    auxiliary mine_labels / mine_valid maps."""
    import train_rl  # noqa: WPS433  (reference script, imported as a module)

    class Stub(torch.nn.Module):
        def __init__(self, HW):
            super().__init__()
            self.HW = HW
            self.g = torch.Generator().manual_seed(5)

        def forward(self, obs, return_mine=False):
            B = obs.shape[0]
            logits = torch.randn((B, self.HW), generator=self.g)
            value = torch.randn((B,), generator=self.g)
            if return_mine:
                return logits, value, torch.zeros((B, 1, obs.shape[-2], obs.shape[-1]))
            return logits, value

    cfg = EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    N, T = 24, 48
    torch.manual_seed(11)
    with LayoutRecorder() as rec:
        vec = VecMinesweeper(N, cfg, seed=3)
        index_of = {id(e): i for i, e in enumerate(vec.envs)}
        # step-resolved layout log: wrap vec.step to know the step index of each placement
        place_t, place_i, place_bits = [], [], []
        tcount = [0]
        orig_step = vec.step

        def step_logged(actions):
            r = orig_step(actions)
            for i, m in rec.drain(index_of):
                place_t.append(tcount[0]); place_i.append(i); place_bits.append(m.reshape(-1))
            tcount[0] += 1
            return r

        vec.step = step_logged
        buf, aux = train_rl.collect_rollout(vec, Stub(256), T, torch.device("cpu"),
                                            aux_mine_weight=0.05, aux_mine_calib_weight=0.01)
    last_values = aux["last_values"].detach().float()
    buf.compute_gae(last_values, gamma=0.995, lam=0.95)
    out = dict(
        H=16, W=16, mine_count=40, safe=1, win_reward=1.0, loss_reward=-1.0, step_penalty=1e-4,
        N=N, T=T,
        actions=buf.actions.view(T, N).numpy(),
        place_t=np.array(place_t, np.int32), place_i=np.array(place_i, np.int32),
        place_bits=np.packbits(np.stack(place_bits).astype(np.uint8), axis=1, bitorder="little"),
        obs=pack(buf.obs.numpy()).reshape(T, N, -1), mask=pack(buf.action_mask.numpy()).reshape(T, N, -1),
        rewards=buf.rewards.view(T, N).numpy(), dones=buf.dones.view(T, N).numpy(),
        values=buf.values.view(T, N).numpy(), logp=buf.logp.view(T, N).numpy(),
        mine_labels=pack(buf.mine_labels.numpy()).reshape(T, N, -1),
        mine_valid=pack(buf.mine_valid.numpy()).reshape(T, N, -1),
        last_values=last_values.numpy(),
        adv=buf.advantages.view(T, N).numpy(), ret=buf.returns.view(T, N).numpy(),
    )
    np.savez_compressed(os.path.join(OUT, "rollout_16x16x40.npz"), **out)
    print(f"rollout: N={N} T={T} dones={int(buf.dones.sum())} placements={len(place_t)}")


def run_late_start_stats():
    """Late-start curriculum (env.py:416-466) draws from ONE sequential generator shared by all envs,
    so it can only be pinned statistically: summary statistics of the reference after reset() and
    after 40 random-valid steps, for three configurations."""
    import json
    out = {}
    cases = {
        "16x16x40_p0.7": (EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4),
                          dict(prob=0.7, min_hidden=5, max_hidden=30, max_attempts=3), 2048),
        "8x8x10_p1.0": (EnvConfig(), dict(prob=1.0, min_hidden=2, max_hidden=6), 4096),
        "4x4x3_p0.9": (EnvConfig(H=4, W=4, mine_count=3), dict(prob=0.9, min_hidden=1, max_hidden=2, max_attempts=2,
                                                                  max_extra_steps=6), 4096),
    }
    for name, (cfg, ls, N) in cases.items():
        HW = cfg.H * cfg.W
        safe = HW - cfg.mine_count
        vec = VecMinesweeper(N, cfg, seed=11, late_start_cfg=ls, late_start_seed=12)
        b = vec.reset()

        def stats():
            f = np.array([e.first_click_done for e in vec.envs])
            hid = np.array([safe - int((e.revealed & ~e.mine_mask).sum()) for e in vec.envs])
            sc = np.array([e.step_count for e in vec.envs])
            return {"frac_started": float(f.mean()), "mean_safe_hidden_started": float(hid[f].mean()) if f.any() else 0.0,
                    "max_safe_hidden_started": int(hid[f].max()) if f.any() else 0,
                    "mean_step_count_started": float(sc[f].mean()) if f.any() else 0.0}

        after_reset = stats()
        rng = np.random.default_rng(1)
        mask = b["action_mask"]
        wins = losses = 0
        for t in range(40):
            s_ = rng.random(mask.shape); s_[~mask] = -1
            b, r, d, info = vec.step(s_.argmax(1).astype(np.int32))
            mask = b["action_mask"]
            wins += sum(o == "win" for o in info["outcome"]); losses += sum(o == "loss" for o in info["outcome"])
        out[name] = {"cfg": dict(H=cfg.H, W=cfg.W, mine_count=cfg.mine_count), "late_start": ls, "N": N,
                     "after_reset": after_reset, "after_40_steps": stats(),
                     "win_rate_random_play": wins / max(1, wins + losses)}
        print("late_start", name, out[name]["after_reset"], out[name]["win_rate_random_play"])
    json.dump(out, open(os.path.join(OUT, "late_start_stats.json"), "w"), indent=1)


def run_forced_subset_case():
    """rules.analyze_forced_modules (rules.py:206-259) on mid-game states: late start + random valid
    play give boards with long frontiers; every env is analysed at several time points."""
    from minesweeper.rules import analyze_forced_modules
    out = {"names": []}
    cases = {
        "16x16x40": (EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4), dict(prob=0.8, min_hidden=20, max_hidden=150), 48),
        "8x8x10": (EnvConfig(), dict(prob=0.8, min_hidden=4, max_hidden=40), 64),
        "16x30x99": (EnvConfig(H=16, W=30, mine_count=99), dict(prob=0.8, min_hidden=40, max_hidden=300), 24),
        "5x7x6": (EnvConfig(H=5, W=7, mine_count=6), dict(prob=0.5, min_hidden=3, max_hidden=20), 64),
    }
    for name, (cfg, ls, N) in cases.items():
        HW = cfg.H * cfg.W
        vec = VecMinesweeper(N, cfg, seed=21, late_start_cfg=ls, late_start_seed=22)
        b = vec.reset()
        rng = np.random.default_rng(2)
        revs, mines, subs = [], [], []
        hits = 0
        for t in range(6):
            for e in vec.envs:
                sr = analyze_forced_modules(e)["subset_reveal"]
                m = np.zeros(HW, bool); m[list(sr)] = True
                revs.append(e.revealed.reshape(-1).copy()); mines.append(e.mine_mask.reshape(-1).copy()); subs.append(m)
                hits += len(sr)
            mask = b["action_mask"]
            s_ = rng.random(mask.shape); s_[~mask] = -1
            b, _, _, _ = vec.step(s_.argmax(1).astype(np.int32))
        out["names"].append(name)
        out[f"{name}_cfg"] = np.array([cfg.H, cfg.W, cfg.mine_count])
        for key, arr in (("rev", revs), ("mine", mines), ("subset", subs)):
            out[f"{name}_{key}"] = np.packbits(np.stack(arr).astype(np.uint8), axis=1, bitorder="little")
        print(f"forced_subset {name}: {len(revs)} states, {hits} subset_reveal cells")
    out["names"] = np.array(out["names"])
    np.savez_compressed(os.path.join(OUT, "forced_subset.npz"), **out)


def run_avoidability_case():
    """avoidability.analyze_avoidability (avoidability.py:145-394) on mid-game states (late start +
    random valid play), with the chosen cell eval.py:352 would pass: the action about to be played.
    States where the reference's search needs more than a few seconds are skipped (and counted)."""
    import time
    from minesweeper.avoidability import analyze_avoidability
    out = {"names": []}
    cases = {
        "16x16x40": (EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4), dict(prob=0.8, min_hidden=20, max_hidden=150), 40, 6),
        "8x8x10": (EnvConfig(), dict(prob=0.8, min_hidden=4, max_hidden=40), 64, 6),
        "16x30x99": (EnvConfig(H=16, W=30, mine_count=99), dict(prob=0.8, min_hidden=40, max_hidden=300), 16, 5),
        "5x7x6": (EnvConfig(H=5, W=7, mine_count=6), dict(prob=0.5, min_hidden=3, max_hidden=20), 64, 6),
        "9x9x30_dense": (EnvConfig(H=9, W=9, mine_count=30), dict(prob=0.7, min_hidden=5, max_hidden=40), 48, 6),
        "16x16x90_dense": (EnvConfig(H=16, W=16, mine_count=90), dict(prob=0.9, min_hidden=10, max_hidden=120), 48, 5),
        "6x6x14_dense": (EnvConfig(H=6, W=6, mine_count=14), dict(prob=0.9, min_hidden=2, max_hidden=16), 64, 6),
    }
    from minesweeper import avoidability as ref_av
    searches = [0]
    orig_feasible = ref_av._ConstraintSolver.is_feasible

    def counting_feasible(self, forced=None):
        searches[0] += 1
        return orig_feasible(self, forced)

    ref_av._ConstraintSolver.is_feasible = counting_feasible      # wrapped, not edited: counts exact searches
    for name, (cfg, ls, N, steps) in cases.items():
        HW = cfg.H * cfg.W
        vec = VecMinesweeper(N, cfg, seed=31, late_start_cfg=ls, late_start_seed=32)
        b = vec.reset()
        rng = np.random.default_rng(3)
        rows = {k: [] for k in ("rev", "mine", "fcd", "chosen", "avoidable", "safe", "ncomp", "sizes", "cfs", "ccs")}
        slow = 0
        searched_states = search_found = 0
        for t in range(steps):
            mask = b["action_mask"]
            s_ = rng.random(mask.shape); s_[~mask] = -1
            actions = s_.argmax(1).astype(np.int32)
            for i, e in enumerate(vec.envs):
                # every third query asks about an arbitrary cell (also revealed / non-frontier ones), one in
                # seven passes chosen_cell=None
                q = (i + t) % 21
                chosen = None if q % 7 == 6 else (int(rng.integers(0, HW)) if q % 3 == 2 else int(actions[i]))
                t0 = time.perf_counter()
                before = searches[0]
                res = analyze_avoidability(e, chosen)
                slow += time.perf_counter() - t0 > 5.0
                if searches[0] > before:            # the unit/subset rules found nothing: exact search ran
                    searched_states += 1
                    search_found += bool(res.forced_safe_cells)
                safe = np.zeros(HW, bool); safe[list(res.forced_safe_cells)] = True
                sizes = np.zeros(HW, np.int16); sizes[: len(res.component_sizes)] = res.component_sizes
                rows["rev"].append(e.revealed.reshape(-1).copy()); rows["mine"].append(e.mine_mask.reshape(-1).copy())
                rows["fcd"].append(bool(e.first_click_done)); rows["chosen"].append(-1 if chosen is None else chosen)
                rows["avoidable"].append(bool(res.avoidable)); rows["safe"].append(safe)
                rows["ncomp"].append(len(res.component_sizes)); rows["sizes"].append(sizes)
                rows["cfs"].append(bool(res.chosen_is_forced_safe))
                rows["ccs"].append(-1 if res.chosen_component_size is None else int(res.chosen_component_size))
            b, _, _, _ = vec.step(actions)
        out["names"].append(name)
        out[f"{name}_cfg"] = np.array([cfg.H, cfg.W, cfg.mine_count])
        for key in ("rev", "mine", "safe"):
            out[f"{name}_{key}"] = np.packbits(np.stack(rows[key]).astype(np.uint8), axis=1, bitorder="little")
        for key, dt in (("fcd", np.uint8), ("chosen", np.int32), ("avoidable", np.uint8), ("ncomp", np.int32),
                        ("sizes", np.int16), ("cfs", np.uint8), ("ccs", np.int32)):
            out[f"{name}_{key}"] = np.asarray(rows[key]).astype(dt)
        nq = len(rows["fcd"])
        print(f"avoidability {name}: {nq} states, avoidable {int(np.sum(rows['avoidable']))}, "
              f"forced-safe cells {int(np.stack(rows['safe']).sum())}, components {int(np.sum(rows['ncomp']))}, "
              f"largest {int(np.stack(rows['sizes']).max())}, states that reached the exact search: {searched_states} "
              f"(it found safe cells in {search_found}), searches > 5 s: {slow}")
    ref_av._ConstraintSolver.is_feasible = orig_feasible
    out["names"] = np.array(out["names"])
    np.savez_compressed(os.path.join(OUT, "avoidability.npz"), **out)


def run_eval_case():
    """BASELINE configs[0] (C1) as a fixture: the UNMODIFIED reference eval.evaluate_vec (eval.py:265-511)
    driven by tests/parity.scripted_policy (exact integer outputs on any device) on the reference env;
    records its metric dict and, per env, every mine layout in the order it was drawn."""
    import json
    sys.path.insert(0, os.path.dirname(OUT))
    import parity
    import eval as ref_eval
    out = {}
    cases = {
        "8x8x10": (EnvConfig(), 24, 96),
        "16x16x40": (EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4), 16, 40),
    }
    for name, (cfg, num_envs, episodes) in cases.items():
        model = parity.scripted_policy(cfg.H, cfg.W, seed=5)
        with LayoutRecorder() as rec:
            created = []
            orig_init = MinesweeperEnv.__init__

            def tracking_init(self, *a, **k):
                orig_init(self, *a, **k)
                created.append(self)

            MinesweeperEnv.__init__ = tracking_init          # wrapped, not edited: env identity -> index
            try:
                metrics = ref_eval.evaluate_vec(model, cfg, episodes=episodes, seed=7, num_envs=num_envs)
            finally:
                MinesweeperEnv.__init__ = orig_init
        index = {id(e): i for i, e in enumerate(created)}
        env_i = np.array([index[id(e)] for e, _ in rec.log], np.int32)
        layouts = pack(np.stack([m.reshape(-1) for _, m in rec.log]))
        out[f"{name}_cfg"] = np.array([cfg.H, cfg.W, cfg.mine_count, num_envs, episodes, 7, 5])
        out[f"{name}_layout_env"] = env_i
        out[f"{name}_layout_bits"] = layouts
        out[f"{name}_metrics"] = np.array(json.dumps(metrics))
        print(f"eval {name}: {len(env_i)} layouts;", {k: round(v, 4) for k, v in metrics.items()})
    out["names"] = np.array(list(cases))
    np.savez_compressed(os.path.join(OUT, "c1_eval.npz"), **out)


def main():
    if "--eval-only" in sys.argv:
        run_eval_case()
        return
    if "--avoid-only" in sys.argv:
        run_avoidability_case()
        return
    if "--subset-only" in sys.argv:
        run_forced_subset_case()
        return
    if "--late-only" in sys.argv:
        run_late_start_stats()
        return
    std = dict(guarantee_safe_neighborhood=True, step_penalty=1e-4)
    run_env_case("16x16x40_valid", EnvConfig(H=16, W=16, mine_count=40, **std), 48, 96, "valid", 0)
    run_env_case("16x16x40_any", EnvConfig(H=16, W=16, mine_count=40, **std), 48, 96, "any", 1)
    run_env_case("16x16x40_win", EnvConfig(H=16, W=16, mine_count=40, **std), 8, 220, "win", 2)
    run_env_case("16x30x99_valid", EnvConfig(H=16, W=30, mine_count=99, **std), 32, 64, "valid", 3)
    run_env_case("16x30x99_win", EnvConfig(H=16, W=30, mine_count=99, **std), 4, 260, "win", 4)
    run_env_case("8x8x10_default", EnvConfig(), 32, 64, "valid", 5)
    run_env_case("4x4x8_fallback", EnvConfig(H=4, W=4, mine_count=8), 32, 48, "any", 6)
    run_env_case("5x7x6_nosafe", EnvConfig(H=5, W=7, mine_count=6, guarantee_safe_neighborhood=False,
                                           win_reward=2.5, loss_reward=-0.75, step_penalty=0.01),
                 32, 64, "valid", 7)
    run_env_case("32x32x150_valid", EnvConfig(H=32, W=32, mine_count=150), 8, 48, "valid", 8)
    run_env_case("30x16x99_win", EnvConfig(H=30, W=16, mine_count=99), 4, 260, "win", 9)
    run_floodfill_case()
    run_gae_cases()
    run_rollout_case()
    run_late_start_stats()
    run_forced_subset_case()
    run_avoidability_case()
    run_eval_case()


if __name__ == "__main__":
    main()
