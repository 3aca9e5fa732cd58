"""Poor man's memcheck (compute-sanitizer is closed on the GPU pool): every output of every kernel
is a window inside a larger canary-filled allocation; after the launch the canaries on both sides
must be untouched.  Odd shapes and tails on purpose."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PAD = 4096          # bytes of canary on each side (multiple of 32 so windows stay aligned)


def _window(torch, shape, dtype, fill=0x5A):
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    raw = torch.full((PAD + n + PAD,), fill, dtype=torch.uint8, device="cuda")
    view = raw[PAD:PAD + n].view(dtype).view(*shape)
    return raw, view, n


def _intact(raw, n, fill=0x5A):
    return bool((raw[:PAD] == fill).all()) and bool((raw[PAD + n:] == fill).all())


@pytest.mark.parametrize("H,W,M,N", [(16, 16, 40, 37), (16, 30, 99, 19), (5, 7, 6, 33), (32, 32, 200, 5), (3, 4, 2, 65)])
def test_env_outputs_stay_inside_their_windows(H, W, M, N):
    import torch
    import minesweeper_ppo_b200 as m
    cfg = m.EnvConfig(H=H, W=W, mine_count=M, step_penalty=1e-4)
    vec = m.VecMinesweeper(N, cfg, seed=1, api="torch", aux_maps=True)
    wins = {
        "obs": _window(torch, (N, 10, H, W), torch.float32), "mask": _window(torch, (N, H * W), torch.bool),
        "lab": _window(torch, (N, H, W), torch.float32), "val": _window(torch, (N, H, W), torch.bool),
        "rew": _window(torch, (N,), torch.float32), "done": _window(torch, (N,), torch.bool),
    }
    out = m.StepOut(obs=wins["obs"][1], action_mask=wins["mask"][1], rewards=wins["rew"][1], dones=wins["done"][1],
                    mine_labels=wins["lab"][1], mine_valid=wins["val"][1])
    vec.reset(out=out)
    for t in range(10):
        vec.step(vec.random_actions(t, valid_only=bool(t % 3)), out=out)
        vec.step_random(50 + t, out=out)
    vec.encode(out=out)
    torch.cuda.synchronize()
    for k, (raw, view, n) in wins.items():
        assert _intact(raw, n), f"{k} window overrun at {H}x{W} N={N}"
    assert bool(((wins["obs"][1] == 0) | (wins["obs"][1] == 1)).all())
    # state tensors too: wrap them in windows by swapping the env's storage
    comp = m.CompactRolloutBuffer(vec, 3, aux_maps=True)
    for t in range(3):
        comp.snapshot(t)
    rows = torch.randperm(3 * N, device="cuda")[: N + 1]
    got = comp.gather_obs(rows)
    assert got.obs.shape[0] == N + 1


# ragged N: the plain-load kernel; N % 16 == 0: the TMA kernel, whose stores the tensor map clips (ragged slices and
# column boxes, both launch shapes)
@pytest.mark.parametrize("T,N", [(1, 1), (129, 33), (5, 70), (128, 31), (70, 48), (161, 16), (1, 16), (33, 14240)])
def test_gae_outputs_stay_inside_their_windows(T, N):
    import torch
    import minesweeper_ppo_b200 as m
    buf = m.RolloutBuffer(N, T, (1, 1, 1), 1, torch.device("cuda"))
    ra, adv, na = _window(torch, (T * N,), torch.float32)
    rr, ret, nr = _window(torch, (T * N,), torch.float32)
    buf.advantages, buf.returns = adv, ret
    buf.rewards.normal_(); buf.values.normal_()
    buf.compute_gae(torch.randn(N, device="cuda"))
    torch.cuda.synchronize()
    assert buf.advantages.data_ptr() == adv.data_ptr()
    assert _intact(ra, na) and _intact(rr, nr)
    assert bool(torch.isfinite(adv).all())


@pytest.mark.parametrize("A,n", [(256, 37), (35, 9), (1024, 3), (480, 65)])
def test_sampler_and_gn_outputs_stay_inside_their_windows(A, n):
    import torch
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200 import _lib
    r64, a64, n64 = _window(torch, (n,), torch.int64)
    r32, a32, n32 = _window(torch, (n,), torch.int32)
    rlp, lp, nlp = _window(torch, (n,), torch.float32)
    logits = torch.randn(n, A, device="cuda").half()
    mask = torch.rand(n, A, device="cuda") < 0.5
    mask[:, 0] = True
    m.masked_sample(logits, mask, seed=1, step_index=0, actions64=a64, actions32=a32, logp=lp)
    torch.cuda.synchronize()
    assert _intact(r64, n64) and _intact(r32, n32) and _intact(rlp, nlp)
    assert int(a64.min()) >= 0 and int(a64.max()) < A
    # msw_gn_act with windows for both outputs
    L = _lib.load()
    C, G, H, W = 32, 2, 5, 7
    x = torch.randn(n, C, H, W, device="cuda").half().contiguous(memory_format=torch.channels_last)
    ry16, y16, ny16 = _window(torch, (n, H, W, C), torch.float16)
    ry32, y32, ny32 = _window(torch, (n, H, W, C), torch.float32)
    rpl, pool, npl = _window(torch, (n, C), torch.float32)
    gn = torch.nn.GroupNorm(G, C).cuda()
    rc = L.msw_gn_act(x.data_ptr(), None, None, gn.weight.data_ptr(), gn.bias.data_ptr(), y16.data_ptr(), y32.data_ptr(),
                      n, H * W, C, G, 1e-5, 1, 0.0, 0, 0, None, None, None, None, pool.data_ptr(), 0,
                      torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert _intact(ry16, ny16) and _intact(ry32, ny32) and _intact(rpl, npl)
    want = torch.relu(gn(x.float())).permute(0, 2, 3, 1)
    assert float((y32 - want.detach()).abs().max()) < 1e-4
    assert float((pool - want.detach().mean(dim=(1, 2))).abs().max()) < 1e-5
