"""RolloutBuffer.compute_gae on CUDA (msw_gae) vs fixtures from the reference's own
compute_gae (buffers.py:78-94) and vs the oracle at C3 size.  Contract: <= 1e-6 relative
(BASELINE.json); the kernel is additionally expected to be bit-exact."""
import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu
REL_TOL = 1e-6          # north_star: "GAE advantages match the reference within 1e-6 relative in fp32"


def _run(torch, rewards, values, dones, last_values, gamma, lam):
    import minesweeper_ppo_b200 as m
    T, N = rewards.shape
    dev = torch.device("cuda")
    buf = m.RolloutBuffer(N, T, (1, 1, 1), 1, dev)
    buf.rewards.copy_(torch.from_numpy(rewards.reshape(-1)))
    buf.values.copy_(torch.from_numpy(values.reshape(-1)))
    buf.dones.copy_(torch.from_numpy(dones.reshape(-1)))
    buf.compute_gae(last_values.to(dev), gamma=gamma, lam=lam)
    return buf.advantages.view(T, N).cpu().numpy(), buf.returns.view(T, N).cpu().numpy()


def _close(a, b):
    denom = np.maximum(np.abs(b), 1e-30)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0


def test_gae_matches_reference_fixtures():
    import torch
    g = P.load("gae")
    for name in g["names"]:
        gamma, lam = (float(x) for x in g[f"{name}_gamma_lam"])
        lv = torch.from_numpy(g[f"{name}_last_values"])
        if bool(g[f"{name}_lv_is_fp16"]):
            lv = lv.half()            # exact: the fixture stores float(fp16 value)
        adv, ret = _run(torch, g[f"{name}_rewards"], g[f"{name}_values"], g[f"{name}_dones"], lv, gamma, lam)
        assert _close(adv, g[f"{name}_adv"]) <= REL_TOL and _close(ret, g[f"{name}_ret"]) <= REL_TOL, name
        P.assert_bits_equal(adv, g[f"{name}_adv"], f"gae {name} adv (bit-exact)")
        P.assert_bits_equal(ret, g[f"{name}_ret"], f"gae {name} ret (bit-exact)")


# N % 16 == 0 takes the pipelined TMA kernel: more slices than ring stages (T > 128), a ragged newest slice
# (T % 32 != 0), a ragged last column box (N % 32 != 0), T < one slice, and both launch shapes (the register-column
# walk up to 3 CTAs per SM = 14,208 columns on 148 SMs, the 8-row-batch walk beyond); the other shapes take the
# plain-load kernel
@pytest.mark.parametrize("T,N", [(128, 8192), (64, 128), (300, 1000), (1, 1), (129, 33), (7, 65536), (300, 1024),
                                 (128, 48), (5, 16), (257, 2000), (100, 16384), (161, 32768), (32, 14208), (33, 14240)])
def test_gae_matches_oracle(oracle, T, N):
    import torch
    rng = np.random.default_rng(T * 100003 + N)
    consts = np.array([-1e-4, -1.0 - 1e-4, 1.0 - 1e-4]).astype(np.float32)
    dones = rng.random((T, N)) < 0.15
    rewards = np.where(dones, consts[rng.integers(1, 3, (T, N))], consts[0]).astype(np.float32)
    values = (0.5 * rng.standard_normal((T, N))).astype(np.float32)
    last = (0.5 * rng.standard_normal(N)).astype(np.float32)
    adv, ret = _run(torch, rewards, values, dones, torch.from_numpy(last), 0.995, 0.95)
    a0, r0 = oracle.gae(rewards, values, dones, last, 0.995, 0.95)
    assert _close(adv, a0) <= REL_TOL and _close(ret, r0) <= REL_TOL
    P.assert_bits_equal(adv, a0, "adv")
    P.assert_bits_equal(ret, r0, "ret")


def test_rollout_fixture_through_buffer_slots():
    """The real collect_rollout trace (train_rl.py:155-289): env kernel writes straight into
    RolloutBuffer slots, then GAE -- every buffer field equals the reference's buffer."""
    import torch
    import minesweeper_ppo_b200 as m
    g = P.load("rollout_16x16x40")
    cfg = P.cfg_of(g)
    N, T, HW = int(g["N"]), int(g["T"]), cfg.H * cfg.W
    v = m.VecMinesweeper(N, m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4), api="torch", aux_maps=True)
    buf = m.RolloutBuffer(N, T, (10, 16, 16), 256, v.device, aux_maps=True)
    scratch = v._alloc_encode()
    v.reset(out=buf.slot(0))
    for t in range(T):
        v.inject_layouts(*P.injections(g, t, N, HW))
        nxt = buf.slot(t + 1) if t + 1 < T else scratch
        cur = buf.slot(t)
        v.step(torch.from_numpy(g["actions"][t]).to(v.device),
               out=m.StepOut(obs=nxt.obs, action_mask=nxt.action_mask, rewards=cur.rewards, dones=cur.dones,
                             mine_labels=nxt.mine_labels, mine_valid=nxt.mine_valid), want_infos=False)
    buf.values.copy_(torch.from_numpy(g["values"].reshape(-1)))
    buf.compute_gae(torch.from_numpy(g["last_values"]).to(v.device), 0.995, 0.95)
    c = lambda t_: t_.cpu().numpy()
    P.assert_bits_equal(c(buf.obs).reshape(T, N, -1), P.unpack(g["obs"], 10 * HW).astype(np.float32), "buffer.obs")
    P.assert_bits_equal(c(buf.action_mask).reshape(T, N, -1), P.unpack(g["mask"], HW), "buffer.action_mask")
    P.assert_bits_equal(c(buf.rewards).reshape(T, N), g["rewards"], "buffer.rewards")
    P.assert_bits_equal(c(buf.dones).reshape(T, N), g["dones"], "buffer.dones")
    P.assert_bits_equal(c(buf.mine_labels).reshape(T, N, -1), P.unpack(g["mine_labels"], HW).astype(np.float32), "buffer.mine_labels")
    P.assert_bits_equal(c(buf.mine_valid).reshape(T, N, -1), P.unpack(g["mine_valid"], HW), "buffer.mine_valid")
    P.assert_bits_equal(c(buf.advantages).reshape(T, N), g["adv"], "buffer.advantages")
    P.assert_bits_equal(c(buf.returns).reshape(T, N), g["ret"], "buffer.returns")
