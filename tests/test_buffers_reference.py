"""CPU: RolloutBuffer.add and the dense get_minibatches against the reference's own RolloutBuffer
(minesweeper/buffers.py:38-76, 96-116; imported unmodified from baseline/_ref) on the recorded rollout fixture.
Both run on CPU tensors here (only compute_gae needs the GPU)."""
import numpy as np
import pytest
import torch

import parity as P
import reference_live as RL

pytestmark = pytest.mark.skipif(not RL.available(), reason="baseline/_ref not installed (tools/install_reference.sh)")

FIELDS = ("obs", "action_mask", "actions", "logp", "rewards", "dones", "values", "advantages", "returns",
          "mine_labels", "mine_valid")


def _fill(buf, steps, N, H, W, with_aux, with_valid, seed):
    g = torch.Generator().manual_seed(seed)
    for t in range(steps):
        obs = (torch.rand((N, 10, H, W), generator=g) < 0.3).float()
        mask = torch.rand((N, H * W), generator=g) < 0.6
        actions = torch.randint(0, H * W, (N,), generator=g)
        logp = -torch.rand((N,), generator=g)
        rewards = torch.randn((N,), generator=g)
        dones = torch.rand((N,), generator=g) < 0.2
        values = torch.randn((N,), generator=g)
        kw = {}
        if with_aux and t >= 1:                 # the reference allocates the aux maps lazily, at the first labelled step
            kw["mine_labels"] = (torch.rand((N, H, W), generator=g) < 0.15).float()
            if with_valid:
                kw["mine_valid"] = torch.rand((N, H, W), generator=g) < 0.5
        buf.add(obs, mask, actions, logp, rewards, dones, values, **kw)


@pytest.mark.parametrize("with_aux,with_valid", [(False, False), (True, True), (True, False)])
def test_add_and_minibatches_match_reference(with_aux, with_valid):
    import minesweeper_ppo_b200 as m
    ref_buffers = RL.load()["buffers"]
    N, T, H, W = 12, 5, 6, 7
    ours = m.RolloutBuffer(N, T, (10, H, W), H * W, torch.device("cpu"))
    ref = ref_buffers.RolloutBuffer(num_envs=N, steps=T, obs_shape=(10, H, W), action_dim=H * W, device=torch.device("cpu"))
    _fill(ours, T, N, H, W, with_aux, with_valid, 3)
    _fill(ref, T, N, H, W, with_aux, with_valid, 3)
    assert ours._t == ref._t == T
    for f in FIELDS:
        a, b = getattr(ours, f), getattr(ref, f)
        assert (a is None) == (b is None), f
        if a is not None:
            assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), f
    # GAE fields by the reference's own loop on both, so the minibatches carry the same advantages / returns
    lv = torch.randn((N,), generator=torch.Generator().manual_seed(9))
    ref.compute_gae(lv, 0.995, 0.95)
    ours.advantages, ours.returns = ref.advantages.clone(), ref.returns.clone()
    for bs in (16, 60, 7):
        torch.manual_seed(11)
        got = list(ours.get_minibatches(bs))
        torch.manual_seed(11)
        want = list(ref.get_minibatches(bs))
        assert len(got) == len(want)
        for x, y in zip(got, want):
            names_x = sorted(k for k in vars(x) if not k.startswith("__"))
            names_y = sorted(k for k in vars(y) if not k.startswith("__"))
            assert names_x == names_y
            for k in names_x:
                assert torch.equal(getattr(x, k), getattr(y, k)), k


def test_add_on_the_recorded_rollout_fixture():
    """The real collect_rollout trace (tests/golden/rollout_16x16x40.npz, recorded from train_rl.py:155-289):
    feeding its per-step tensors through `add` -- ours and the reference's -- builds identical buffers, and
    the dense minibatches drawn from them are identical too."""
    import minesweeper_ppo_b200 as m
    ref_buffers = RL.load()["buffers"]
    g = P.load("rollout_16x16x40")
    N, T, H, W = int(g["N"]), int(g["T"]), int(g["H"]), int(g["W"])
    HW = H * W
    obs = P.unpack(g["obs"], 10 * HW).reshape(T, N, 10, H, W).astype(np.float32)
    mask = P.unpack(g["mask"], HW).reshape(T, N, HW)
    lab = P.unpack(g["mine_labels"], HW).reshape(T, N, H, W).astype(np.float32)
    val = P.unpack(g["mine_valid"], HW).reshape(T, N, H, W)
    ours = m.RolloutBuffer(N, T, (10, H, W), HW, torch.device("cpu"))
    ref = ref_buffers.RolloutBuffer(num_envs=N, steps=T, obs_shape=(10, H, W), action_dim=HW, device=torch.device("cpu"))
    for buf in (ours, ref):
        for t in range(T):
            buf.add(torch.from_numpy(obs[t]), torch.from_numpy(mask[t]), torch.from_numpy(g["actions"][t]),
                    torch.from_numpy(g["logp"][t]), torch.from_numpy(g["rewards"][t]), torch.from_numpy(g["dones"][t]),
                    torch.from_numpy(g["values"][t]), mine_labels=torch.from_numpy(lab[t]), mine_valid=torch.from_numpy(val[t]))
    ref.compute_gae(torch.from_numpy(g["last_values"]), 0.995, 0.95)
    assert np.array_equal(ref.advantages.view(T, N).numpy().view(np.uint32), g["adv"].view(np.uint32))   # the fixture's own GAE
    ours.advantages, ours.returns = ref.advantages.clone(), ref.returns.clone()
    for f in FIELDS:
        assert torch.equal(getattr(ours, f), getattr(ref, f)), f
    torch.manual_seed(5)
    got = list(ours.get_minibatches(256))
    torch.manual_seed(5)
    want = list(ref.get_minibatches(256))
    for x, y in zip(got, want):
        for k in (k for k in vars(y) if not k.startswith("__")):
            assert torch.equal(getattr(x, k), getattr(y, k)), k
