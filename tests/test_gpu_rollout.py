"""Fused masked sampler (msw_masked_sample) and the device-resident collector (rollout.py),
the rows of SURVEY section 8 next to the env path (a12, f1).  Needs a B200."""
import numpy as np
import pytest

import parity as P

pytestmark = pytest.mark.gpu
LOGP_TOL = 2e-5      # fp32 log-softmax vs torch's fp32 log_softmax of the same masked logits


@pytest.mark.parametrize("dtype_name", ["float32", "float16", "bfloat16"])
@pytest.mark.parametrize("A", [256, 480, 64, 35, 1024])
def test_masked_sample_logp_and_validity(dtype_name, A):
    import torch
    import minesweeper_ppo_b200 as m
    dt = getattr(torch, dtype_name)
    g = torch.Generator(device="cuda").manual_seed(A)
    n = 4096
    logits = (3.0 * torch.randn((n, A), device="cuda", generator=g)).to(dt)
    mask = torch.rand((n, A), device="cuda", generator=g) < 0.4
    mask[:, 0] |= ~mask.any(1)                                   # at least one legal action per row
    mask[5] = False; mask[5, A - 1] = True                       # single legal action
    a64, a32, logp = m.masked_sample(logits, mask, seed=7, step_index=3)
    assert torch.equal(a64.to(torch.int32), a32)
    assert bool(mask.gather(1, a64[:, None]).all()), "sampled an illegal action"
    assert int(a64[5]) == A - 1
    neg = -1e9 if dt == torch.float32 else -1e4                  # train_rl.py:229-232
    ref = torch.log_softmax(logits.masked_fill(~mask, neg).float(), dim=-1).gather(1, a64[:, None])[:, 0]
    assert float((logp - ref).abs().max()) <= LOGP_TOL
    # deterministic in (seed, step, row); different step -> different draws
    b64, _, _ = m.masked_sample(logits, mask, seed=7, step_index=3)
    c64, _, _ = m.masked_sample(logits, mask, seed=7, step_index=4)
    assert torch.equal(a64, b64) and not torch.equal(a64, c64)


def test_masked_sample_distribution_chi2():
    import torch
    from scipy import stats
    import minesweeper_ppo_b200 as m
    A, n = 16, 400_000
    row = torch.tensor([0.3, -1.0, 2.0, 0.0, 1.5, -3.0, 0.7, 0.1, 9.0, -0.2, 1.1, 0.4, -0.6, 2.2, 0.9, 0.0])
    mask_row = torch.ones(A, dtype=torch.bool); mask_row[8] = False; mask_row[3] = False
    logits = row.cuda().repeat(n, 1).half()
    mask = mask_row.cuda().repeat(n, 1).contiguous()
    a64, _, _ = m.masked_sample(logits, mask, seed=123, step_index=0)
    counts = torch.bincount(a64, minlength=A).cpu().numpy().astype(np.float64)
    assert counts[8] == 0 and counts[3] == 0
    p = torch.softmax(logits[0].float().masked_fill(~mask[0], -1e4), 0).cpu().numpy().astype(np.float64)
    keep = p > 1e-12
    chi2 = ((counts[keep] - n * p[keep]) ** 2 / (n * p[keep])).sum()
    assert stats.chi2.sf(chi2, df=int(keep.sum()) - 1) > 1e-4, chi2


def test_collector_buffer_is_self_consistent(oracle):
    """collect_rollout on device: replay the recorded actions through the CPU oracle (same board
    sampler spec, same seed) and require every buffer field the env produced to match bit for bit."""
    import os
    import torch
    import minesweeper_ppo_b200 as m
    torch.manual_seed(0)
    N, T = 256, 24
    cfg = m.EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    vec = m.VecMinesweeper(N, cfg, seed=11, api="torch")
    model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                          model_cfg=dict(stem_channels=32, blocks=2, dropout=0.0, value_hidden=64)).cuda().eval()
    buf, aux = m.collect_rollout(vec, model, T, torch.device("cuda"), aux_mine_weight=0.05, aux_mine_calib_weight=0.01)
    assert aux["last_values"].shape == (N,) and aux["last_values"].dtype == torch.float16     # train_rl.py:272-277
    assert set(aux["timings"]) >= {"steps", "rollout_total_s"}
    buf.compute_gae(aux["last_values"], 0.995, 0.95)
    c = lambda t_: t_.cpu().numpy()
    acts = c(buf.actions).reshape(T, N)
    ref = oracle.OracleVecEnv(N, cfg, seed=11, nthreads=os.cpu_count() or 1, aux_maps=True)
    b = ref.reset()                   # collect_rollout resets at the start of every rollout (train_rl.py:163)
    obs, mask, lab, val = b["obs"], b["action_mask"], ref.mine_labels, ref.mine_valid
    for t in range(T):
        s = slice(t * N, (t + 1) * N)
        P.assert_bits_equal(c(buf.obs[s]), obs, f"slot {t} obs")
        P.assert_bits_equal(c(buf.action_mask[s]), mask, f"slot {t} mask")
        P.assert_bits_equal(c(buf.mine_labels[s]), lab, f"slot {t} labels")
        P.assert_bits_equal(c(buf.mine_valid[s]), val, f"slot {t} valid")
        assert mask[np.arange(N), acts[t]].all(), "policy sampled a revealed cell"
        bb, r, d, _ = ref.step(acts[t], tensor_infos=True)
        P.assert_bits_equal(c(buf.rewards[s]), r, f"slot {t} rewards")
        P.assert_bits_equal(c(buf.dones[s]), d, f"slot {t} dones")
        obs, mask, lab, val = bb["obs"], bb["action_mask"], ref.mine_labels, ref.mine_valid
    # GAE on the collected buffer == oracle GAE with the fp16-bootstrap rule (buffers.py:88-90)
    lv16 = aux["last_values"]
    T_, N_ = T, N
    r_np, v_np, d_np = c(buf.rewards).reshape(T_, N_), c(buf.values).reshape(T_, N_), c(buf.dones).reshape(T_, N_)
    gv_last = (0.995 * lv16).float().cpu().numpy()              # torch's own fp16 multiply
    # oracle GAE takes last_values and multiplies by gamma in fp32; emulate the prescaled first step by
    # folding it into the reward of the last row: r' = r + gv_last*nnt, last_values = 0
    nnt = 1.0 - d_np[-1].astype(np.float32)
    r2 = r_np.copy(); r2[-1] = (r_np[-1] + gv_last * nnt).astype(np.float32)
    adv, ret = oracle.gae(r2, v_np, d_np, np.zeros(N_, np.float32), 0.995, 0.95)
    P.assert_bits_equal(c(buf.advantages).reshape(T_, N_), adv, "advantages")
    P.assert_bits_equal(c(buf.returns).reshape(T_, N_), ret, "returns")
    # second rollout reuses nothing stale
    col = m.RolloutCollector(vec, T, aux_maps=False)
    b1, _ = col.collect(model)
    a1 = b1.actions.clone()
    b2, _ = col.collect(model)
    assert not torch.equal(a1, b2.actions)


@pytest.mark.parametrize("C,blocks,HW", [(96, 5, (16, 16)), (32, 2, (16, 30)), (128, 1, (8, 8))])
def test_fused_forward_matches_autocast_module(C, blocks, HW):
    """msw_gn_act path vs the unchanged module under fp16 autocast (train_rl.py:222): same rounding
    points, so outputs agree to fp16 resolution.  Dropout off (its RNG stream is not torch's)."""
    import torch
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200.fused_forward import FusedRolloutForward, gn_act
    torch.manual_seed(1)
    H, W = HW
    net = m.build_model("cnn_residual", obs_shape=(10, H, W),
                        model_cfg=dict(stem_channels=C, blocks=blocks, dropout=0.0, value_hidden=64)).cuda()
    with torch.no_grad():                                    # non-trivial affine parameters
        for mod in net.modules():
            if isinstance(mod, torch.nn.GroupNorm):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.3, 0.3)
    x = (torch.rand(64, 10, H, W, device="cuda") < 0.3).float()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ref = net(x, return_mine=True)
    ff = FusedRolloutForward(net)
    got = ff(x, return_mine=True)
    again = ff(x, return_mine=True)
    assert all(torch.equal(u, v) for u, v in zip(got, again)), "fused forward must be bitwise reproducible"
    for a, b, name in zip(got, ref, ("logits", "value", "mine")):
        assert a.shape == b.shape and a.dtype == torch.float16, name
        err = float((a.float() - b.float()).abs().max())
        scale = float(b.float().abs().max()) + 1e-3
        assert err <= 2e-2 * scale, (name, err, scale)
    # the kernel alone vs torch GroupNorm in fp32 on the same fp16 input
    gn = torch.nn.GroupNorm(C // 16, C).cuda()
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5); gn.bias.uniform_(-0.3, 0.3)
    z = torch.randn(32, C, H, W, device="cuda").half().contiguous(memory_format=torch.channels_last)
    r = torch.randn(32, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    y16, y32 = gn_act(z, gn, res32=r, relu=True, want32=True)
    want = torch.relu(gn(z.float()) + r)
    assert float((y32 - want.detach()).abs().max()) <= 2e-5 * (float(want.abs().max()) + 1)     # fp32 path tolerance
    assert torch.equal(y16, y32.half())
    # Dropout2d: whole channels dropped with probability p, survivors scaled by 1/(1-p)
    y16, _ = gn_act(z, gn, relu=True, drop_p=0.25, seed=3, call_id=9)
    base, _ = gn_act(z, gn, relu=True)
    dropped = (y16.float().abs().sum(dim=(2, 3)) == 0) & (base.float().abs().sum(dim=(2, 3)) > 0)
    frac = float(dropped.float().mean())
    assert 0.2 < frac < 0.3, frac
    keep = ~dropped
    ratio = (y16.float().sum(dim=(2, 3))[keep] / base.float().sum(dim=(2, 3))[keep])
    assert float((ratio - 1 / 0.75).abs().max()) < 2e-2


@pytest.mark.parametrize("C,R", [(64, 1000), (96, 128 * 700 + 5), (32, 77), (128, 4096)])
def test_cell_heads_match_torch(C, R):
    """msw_cell_heads vs the same two-layer heads as fp16 torch linears (fp32 accumulation, hidden layer
    rounded to fp16 before the ReLU), including a ragged last tile and more tiles than resident CTAs."""
    import torch
    from minesweeper_ppo_b200.fused_forward import cell_heads
    g = torch.Generator(device="cuda").manual_seed(C + R)
    rows = torch.randn((R, C), device="cuda", generator=g).half()
    w1 = (torch.randn((2 * C, C), device="cuda", generator=g) / C ** 0.5).half()
    b1 = (0.1 * torch.randn((2 * C,), device="cuda", generator=g)).half()
    w2 = (torch.randn((2 * C,), device="cuda", generator=g) / C ** 0.5).half()
    b2 = torch.tensor([0.25, -0.5], device="cuda").half()
    pol, mine = cell_heads(rows, w1, b1, w2, b2)
    pol2, mine2 = cell_heads(rows, w1, b1, w2, b2)
    assert torch.equal(pol, pol2) and torch.equal(mine, mine2)
    hid = torch.relu((rows.float() @ w1.float().t() + b1.float()).half()).float()
    want_p = (hid[:, :C] @ w2[:C].float() + b2[0].float()).half()
    want_m = (hid[:, C:] @ w2[C:].float() + b2[1].float()).half()
    for got, want in ((pol, want_p), (mine, want_m)):
        assert got.shape == (R,) and got.dtype == torch.float16
        # one fp16 ulp of the output scale (summation order differs) plus hidden units that round across an ulp
        assert float((got.float() - want.float()).abs().max()) <= 4e-3 * (float(want.float().abs().max()) + 1.0)


@pytest.mark.parametrize("n,cin", [(1, 96), (5, 96), (300, 96), (3, 16), (301, 16)])
def test_conv3x3_matches_cudnn(n, cin):
    """msw_conv3x3 (tcgen05 implicit GEMM, nine shifted taps) vs cuDNN's fp16 convolution of the same
    operands: both accumulate in fp32 and round once to fp16, so they differ by summation order only."""
    import torch
    import torch.nn.functional as F
    from minesweeper_ppo_b200.fused_forward import conv3x3, conv3x3_taps
    g = torch.Generator(device="cuda").manual_seed(n)
    C = 96
    x = torch.randn((n, cin, 16, 16), device="cuda", generator=g).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((C, cin, 3, 3), device="cuda", generator=g) / (9 * cin) ** 0.5).half()
    got = conv3x3(x, conv3x3_taps(w))
    again = conv3x3(x, conv3x3_taps(w))
    assert torch.equal(got, again)
    want = F.conv2d(x, w.contiguous(memory_format=torch.channels_last), None, padding=1)
    exact = F.conv2d(x.float(), w.float(), None, padding=1)
    err, ref_err = float((got.float() - exact).abs().max()), float((want.float() - exact).abs().max())
    assert got.shape == want.shape and got.dtype == torch.float16
    assert err <= max(2 * ref_err, 4e-3), (err, ref_err)            # as close to the fp32 result as cuDNN is
    # every tap in isolation (one-hot weights) catches a transposed or mis-shifted tap exactly
    for ky in range(3):
        for kx in range(3):
            w1 = torch.zeros((C, cin, 3, 3), device="cuda", dtype=torch.float16)
            w1[:, :, ky, kx] = torch.eye(C, cin, device="cuda", dtype=torch.float16)
            one = conv3x3(x, conv3x3_taps(w1))
            ref = F.conv2d(x, w1.contiguous(memory_format=torch.channels_last), None, padding=1)
            assert torch.equal(one, ref), (ky, kx)


@pytest.mark.parametrize("n,res,drop,pool", [(3, False, 0.0, False), (149, True, 0.0, False), (300, False, 0.25, False),
                                             (7, True, 0.0, False), (5, True, 0.0, True), (301, True, 0.0, True)])
def test_conv3x3_gn_matches_torch_fp32(n, res, drop, pool):
    """msw_conv3x3_gn (conv + GroupNorm [+ fp32 residual] + ReLU [+ Dropout2d] [+ average pool] in one tcgen05
    launch, cnn_residual.py:17-27) against the same expression in plain torch fp32:
        relu(group_norm(fp16(conv_fp32(x16, w16)) + bias) [+ res])
    -- the conv result rounded to fp16 where autocast rounds it, everything after it in fp32.  The only freedom
    is the summation order of the fp32 accumulator (a conv output may land on the other side of an fp16 rounding
    boundary: 1 fp16 ulp of a unit-scale value, scaled by gamma * rstd <= ~2).  The residual stream and the fp32
    output use the kernels' private P8 order (fused_forward.to_p8 / from_p8).  The unfused pair
    msw_conv3x3 -> msw_gn_act is kept as a second, tighter reference (identical fp16 conv rounding)."""
    import torch
    import torch.nn.functional as F
    from minesweeper_ppo_b200.fused_forward import conv3x3, conv3x3_gn, conv3x3_taps, from_p8, gn_act, to_p8
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(n)
    C = 96
    x = torch.randn((n, C, 16, 16), device="cuda", generator=g).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((C, C, 3, 3), device="cuda", generator=g) / (9 * C) ** 0.5).half()
    bias = 0.2 * torch.randn((C,), device="cuda", generator=g)
    norm = torch.nn.GroupNorm(6, C).cuda()
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5); norm.bias.uniform_(-0.3, 0.3)
    r = torch.randn((n, C, 16, 16), device="cuda", generator=g) if res else None
    r_p8 = to_p8(r) if res else None
    assert not res or torch.equal(from_p8(r_p8), r)                      # the layout helpers are inverses
    taps = conv3x3_taps(w)
    kw = dict(res32=r_p8, drop_p=drop, seed=5, call_id=77, sample_id_base=1000)
    got16, got2 = conv3x3_gn(x, taps, norm, bias, want32=not pool, want_pool=pool, **kw)
    again16, again2 = conv3x3_gn(x, taps, norm, bias, want32=not pool, want_pool=pool, **kw)
    assert torch.equal(got16, again16) and torch.equal(got2, again2), "bitwise reproducible"
    # ---- torch fp32 reference
    with torch.no_grad():
        conv16 = F.conv2d(x.float(), w.float(), None, padding=1).half()
        z = F.group_norm(conv16.float() + bias.view(1, C, 1, 1), 6, norm.weight, norm.bias, norm.eps)
        want32 = torch.relu(z + r if res else z)
    scale = float(want32.abs().max()) + 1.0
    if drop > 0:      # Dropout2d: whole channels of a board zeroed, survivors scaled by 1/(1-p); mask = the kernel's stream
        got32 = from_p8(got2)
        dropped = got32.abs().sum(dim=(2, 3)) == 0
        frac = float((dropped & (want32.abs().sum(dim=(2, 3)) > 0)).float().mean())
        assert 0.18 < frac < 0.32, frac
        want32 = torch.where(dropped[:, :, None, None], torch.zeros_like(want32), want32 / (1.0 - drop))
        scale = float(want32.abs().max()) + 1.0
    if pool:
        assert got2.shape == (n, C)
        assert float((got2 - want32.mean(dim=(2, 3))).abs().max()) <= 1e-3 * scale
    else:
        got32 = from_p8(got2)
        assert float((got32 - want32).abs().max()) <= 4e-3 * scale
        assert float((got32 - want32).abs().mean()) <= 2e-5 * scale          # boundary flips are rare
        assert torch.equal(got16, got32.half().contiguous(memory_format=torch.channels_last))
    assert float((got16.float() - want32).abs().max()) <= 5e-3 * scale
    # ---- the unfused pair: same fp16 conv output bit for bit, so only the statistics' reduction order differs
    u16, u32 = gn_act(conv3x3(x, taps), norm, conv_bias=bias, res32=(r.contiguous(memory_format=torch.channels_last) if res else None),
                      drop_p=drop, want32=True, seed=5, call_id=77, sample_id_base=1000)
    uscale = float(u32.abs().max()) + 1.0
    if pool:
        assert float((got2 - u32.mean(dim=(2, 3))).abs().max()) <= 2e-5 * uscale
    else:
        assert float((from_p8(got2) - u32).abs().max()) <= 2e-5 * uscale
        if drop > 0:                                             # the same channels are dropped by both kernels
            assert torch.equal(from_p8(got2).abs().sum(dim=(2, 3)) == 0, u32.abs().sum(dim=(2, 3)) == 0)
    assert float((got16.float() - u16.float()).abs().max()) <= 2e-3 * uscale
    only16, none = conv3x3_gn(x, taps, norm, bias, want32=False, **kw) if not res else (got16, None)
    assert none is None and torch.equal(only16, got16)


def test_dropout_stream_is_keyed_by_global_sample():
    """Dropout2d masks depend on sample_id_base + local index (the global env id), not on the local index:
    a shard starting at 40 draws what rows 40.. of an unsharded call draw (shard invariance of the rollout)."""
    import torch
    from minesweeper_ppo_b200.fused_forward import conv3x3_gn, conv3x3_taps, gn_act
    g = torch.Generator(device="cuda").manual_seed(0)
    C, n = 96, 64
    x = torch.randn((n, C, 16, 16), device="cuda", generator=g).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn((C, C, 3, 3), device="cuda", generator=g) / (9 * C) ** 0.5).half()
    bias = torch.zeros((C,), device="cuda")
    norm = torch.nn.GroupNorm(6, C).cuda()
    taps = conv3x3_taps(w)
    full, _ = conv3x3_gn(x, taps, norm, bias, drop_p=0.3, seed=1, call_id=2, sample_id_base=0)
    part, _ = conv3x3_gn(x[40:].contiguous(memory_format=torch.channels_last), taps, norm, bias, drop_p=0.3, seed=1, call_id=2,
                         sample_id_base=40)
    assert torch.equal(full[40:], part)
    other, _ = conv3x3_gn(x[40:].contiguous(memory_format=torch.channels_last), taps, norm, bias, drop_p=0.3, seed=1, call_id=2,
                          sample_id_base=0)
    assert not torch.equal(full[40:], other)
    z = x
    f2, _ = gn_act(z, norm, drop_p=0.3, seed=1, call_id=2, sample_id_base=0)
    p2, _ = gn_act(z[40:].contiguous(memory_format=torch.channels_last), norm, drop_p=0.3, seed=1, call_id=2, sample_id_base=40)
    assert torch.equal(f2[40:], p2)


def test_gn_act_pooled_output():
    """msw_gn_act pool32 = spatial mean of the fp32 output, with and without writing y32."""
    import torch
    from minesweeper_ppo_b200.fused_forward import gn_act
    torch.manual_seed(5)
    for C, H, W in ((96, 16, 16), (32, 3, 5)):
        gn = torch.nn.GroupNorm(C // 16, C).cuda()
        z = torch.randn(9, C, H, W, device="cuda").half().contiguous(memory_format=torch.channels_last)
        r = torch.randn(9, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
        pool_a = torch.empty((9, C), device="cuda")
        pool_b = torch.empty((9, C), device="cuda")
        y16a, y32 = gn_act(z, gn, res32=r, want32=True, pool32=pool_a)
        y16b, none = gn_act(z, gn, res32=r, want32=False, pool32=pool_b)
        assert none is None and torch.equal(y16a, y16b) and torch.equal(pool_a, pool_b)
        assert float((pool_a - y32.mean(dim=(2, 3))).abs().max()) <= 1e-5 * (float(y32.abs().max()) + 1)


@pytest.mark.parametrize("name", ["8x8x10", "16x16x40"])
def test_eval_matches_reference_metrics(name):
    """BASELINE configs[0] (C1) end to end: tests/golden/c1_eval.npz holds the metric dict the UNMODIFIED
    reference eval.evaluate_vec produced with parity.scripted_policy (exact integer outputs on any device)
    and every mine layout its envs drew.  evaluate_vec of this package, on the CUDA env fed the same
    layouts, must return the same dict: all counts-derived metrics exactly, the two belief metrics to 1e-6
    (sigmoid differs in the last ulp between CPU and GPU)."""
    import json
    import math
    from collections import deque
    import torch
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200.evaluate import evaluate_vec
    import parity as P
    g = P.load("c1_eval")
    H, W, M, num_envs, episodes, seed, pol_seed = (int(x) for x in g[f"{name}_cfg"])
    want = json.loads(str(g[f"{name}_metrics"]))
    layouts = P.unpack(g[f"{name}_layout_bits"], H * W)
    queues = [deque() for _ in range(num_envs)]
    for i, bits in zip(g[f"{name}_layout_env"], layouts):
        queues[int(i)].append(bits)

    class ReplayVec(m.VecMinesweeper):
        """The env under test, drawing the reference's layouts: every env about to make its first click
        of an episode takes the next layout the reference drew for that env."""

        def step(self, actions):
            fresh = self._meta[:, 0].cpu().numpy() == 0
            mine = np.zeros((self.num_envs, self.HW), bool)
            for i in np.flatnonzero(fresh):
                mine[i] = queues[i].popleft()
            self.inject_layouts(mine, fresh)
            return super().step(actions)

    cfg = m.EnvConfig(H=H, W=W, mine_count=M, step_penalty=1e-4)
    model = P.scripted_policy(H, W, seed=pol_seed).cuda()
    got = evaluate_vec(model, cfg, episodes=episodes, seed=seed, num_envs=num_envs, vec_factory=ReplayVec)
    assert all(len(q) == 0 for q in queues), "the replay must consume exactly the layouts the reference drew"
    for k, w in want.items():
        v = got[k]
        if isinstance(w, float) and math.isnan(w):
            assert math.isnan(v), (k, v)
        elif k.startswith("belief_"):
            assert abs(v - w) <= 1e-6, (k, v, w)
        else:
            assert v == w, (k, v, w)


def test_eval_compat_c1():
    """BASELINE.json configs[0]: eval yaml env (16x16x40), 64 envs, 256 episodes, random-init medium
    policy, greedy argmax through the NumPy API and the vec.envs[i] views.  An untrained greedy
    policy loses almost immediately: the survey measured win_rate 0.0 / avg_steps 1.95 on the
    reference; layouts differ (different sampler), so only the regime is
    asserted."""
    import torch
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200.evaluate import evaluate_vec
    torch.manual_seed(0)
    model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                          model_cfg=dict(stem_channels=96, blocks=5, dropout=0.05, value_hidden=256)).cuda()
    cfg = m.EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    r = evaluate_vec(model, cfg, episodes=256, seed=0, num_envs=64)
    assert r["episodes"] == 256 and r["invalid_rate"] == 0.0
    assert r["win_rate"] <= 0.05 and 1.0 <= r["avg_steps"] <= 12.0
    assert 0.0 < r["avg_progress"] < 1.0 and 0.3 < r["belief_auroc"] < 0.7      # untrained belief head ~ chance


@pytest.mark.parametrize("H,W,M", [(16, 16, 40), (16, 30, 99), (9, 9, 10)])
def test_compact_replay_gather_equals_dense_buffer(H, W, M):
    """re-encoding minibatch rows from bitboard snapshots == gathering dense buffer rows."""
    import torch
    import minesweeper_ppo_b200 as m
    N, T = 512, 20
    cfg = m.EnvConfig(H=H, W=W, mine_count=M, step_penalty=1e-4)
    vec = m.VecMinesweeper(N, cfg, seed=4, api="torch", aux_maps=True)
    dense = m.RolloutBuffer(N, T, (10, H, W), H * W, vec.device, aux_maps=True)
    comp = m.CompactRolloutBuffer(vec, T, aux_maps=True)
    scratch = vec._alloc_encode()
    vec.reset(out=dense.slot(0)); comp.snapshot(0)
    for t in range(T):
        nxt = dense.slot(t + 1) if t + 1 < T else scratch
        cur = dense.slot(t)
        vec.step_random(t, out=m.StepOut(obs=nxt.obs, action_mask=nxt.action_mask, rewards=cur.rewards, dones=cur.dones,
                                         mine_labels=nxt.mine_labels, mine_valid=nxt.mine_valid))
        if t + 1 < T:
            comp.snapshot(t + 1)
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = torch.randperm(N * T, device="cuda", generator=g)[:3000]
    got = comp.gather_obs(rows)
    assert torch.equal(got.obs, dense.obs[rows]) and torch.equal(got.action_mask, dense.action_mask[rows])
    assert torch.equal(got.mine_labels, dense.mine_labels[rows]) and torch.equal(got.mine_valid, dense.mine_valid[rows])
    assert float(got.obs.sum()) > 0
    batches = list(comp.get_minibatches(4096))
    assert sum(b.obs.shape[0] for b in batches) == N * T and batches[0].obs.shape[1:] == (10, H, W)


def test_compact_collector_matches_dense_collector():
    import torch
    import minesweeper_ppo_b200 as m
    cfg = m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4)
    torch.manual_seed(0)
    model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                          model_cfg=dict(stem_channels=32, blocks=2, dropout=0.0, value_hidden=64)).cuda().eval()
    outs = []
    for compact in (False, True):
        vec = m.VecMinesweeper(256, cfg, seed=9, api="torch")
        col = m.RolloutCollector(vec, 16, aux_maps=True, sample_seed=5, compact=compact)
        buf, aux = col.collect(model)
        outs.append((buf, aux))
    (d, da), (c, ca) = outs
    for f in ("actions", "logp", "rewards", "dones", "values"):
        assert torch.equal(getattr(d, f), getattr(c, f)), f
    assert torch.equal(da["last_values"], ca["last_values"])
    rows = torch.arange(0, 256 * 16, 7, device="cuda")
    got = c.gather_obs(rows)
    assert torch.equal(got.obs, d.obs[rows]) and torch.equal(got.mine_valid, d.mine_valid[rows])


def test_graph_captured_rollout(oracle):
    """RolloutCollector(graph=True): the whole rollout is one CUDA-graph replay.  Replays must draw
    fresh actions (device-side epoch counter) and stay consistent with the env semantics: replaying
    the recorded actions through the oracle reproduces every env-written buffer field bit for bit."""
    import os
    import torch
    import minesweeper_ppo_b200 as m
    torch.manual_seed(0)
    N, T = 128, 16
    cfg = m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4)
    vec = m.VecMinesweeper(N, cfg, seed=2, api="torch")
    model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                          model_cfg=dict(stem_channels=32, blocks=2, dropout=0.05, value_hidden=64)).cuda()
    col = m.RolloutCollector(vec, T, aux_maps=True, sample_seed=3, graph=True)
    b1, _ = col.collect(model)
    torch.cuda.synchronize()
    a1, v1 = b1.actions.clone(), b1.values.clone()
    ep = vec.state_tensors["meta"][:, 2].cpu().numpy().astype(np.uint32)
    with torch.no_grad():                                   # weights change between rollouts (in-place refresh)
        for p_ in model.parameters():
            p_.mul_(1.01)
    b2, aux = col.collect(model)
    torch.cuda.synchronize()
    assert not torch.equal(a1, b2.actions) and not torch.equal(v1, b2.values)
    c = lambda t_: t_.cpu().numpy()
    ref = oracle.OracleVecEnv(N, cfg, seed=2, nthreads=os.cpu_count() or 1, aux_maps=True)
    ref.episode_idx[:] = ep
    b = ref.reset()
    obs, mask, lab = b["obs"], b["action_mask"], ref.mine_labels
    acts = c(b2.actions).reshape(T, N)
    for t in range(T):
        s = slice(t * N, (t + 1) * N)
        P.assert_bits_equal(c(b2.obs[s]), obs, f"slot {t} obs")
        P.assert_bits_equal(c(b2.action_mask[s]), mask, f"slot {t} mask")
        P.assert_bits_equal(c(b2.mine_labels[s]), lab, f"slot {t} labels")
        assert mask[np.arange(N), acts[t]].all()
        bb, r, d, _ = ref.step(acts[t], tensor_infos=True)
        P.assert_bits_equal(c(b2.rewards[s]), r, f"slot {t} rewards")
        P.assert_bits_equal(c(b2.dones[s]), d, f"slot {t} dones")
        obs, mask, lab = bb["obs"], bb["action_mask"], ref.mine_labels
    assert aux["last_values"].shape == (N,)


def test_gn_act_autograd_matches_torch():
    """msw_gn_act + msw_gn_act_bwd as an autograd Function vs torch autograd of the same expression in
    fp32 (relu(group_norm(x + bias) + res)), with upstream gradients on both outputs."""
    import torch
    import torch.nn.functional as F
    from minesweeper_ppo_b200.fused_train import GnAct
    torch.manual_seed(0)
    for (N, C, G, H, W, with_res) in [(8, 96, 6, 16, 16, True), (5, 32, 2, 5, 7, False), (3, 128, 8, 8, 8, True)]:
        x = torch.randn(N, C, H, W, device="cuda").half().contiguous(memory_format=torch.channels_last).requires_grad_()
        cb = torch.randn(C, device="cuda", requires_grad=True)
        gam = (torch.rand(C, device="cuda") + 0.5).requires_grad_()
        bet = torch.randn(C, device="cuda", requires_grad=True)
        res = torch.randn(N, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_() if with_res else None
        y16, y32 = GnAct.apply(x, cb, gam, bet, res, G, 1e-5, True, 0.0, 0, 0, True)
        u16 = torch.randn_like(y16); u32 = torch.randn_like(y32)
        ((y16.float() * u16.float()).sum() + (y32 * u32).sum()).backward()
        got = [t.grad.clone() for t in (x, cb, gam, bet)] + ([res.grad.clone()] if with_res else [])
        for t in (x, cb, gam, bet) + ((res,) if with_res else ()):
            t.grad = None
        z = F.group_norm(x.float() + cb.view(1, C, 1, 1), G, gam, bet, 1e-5)
        ref = torch.relu(z + res) if with_res else torch.relu(z)
        assert float((y32 - ref).abs().max()) <= 2e-5 * (float(ref.abs().max()) + 1)
        # the kernel sees one upstream gradient per element: g16 (fp16) + g32
        (ref * (u16.float() + u32)).sum().backward()
        want = [t.grad for t in (x, cb, gam, bet)] + ([res.grad] if with_res else [])
        names = ["dx", "dbias", "dgamma", "dbeta", "dres"]
        for a, b, nm in zip(got, want, names):
            err = float((a.float() - b.float()).abs().max())
            scale = float(b.float().abs().max()) + 1e-6
            tol = 4e-3 if nm == "dx" else 2e-4            # dx is rounded to fp16
            assert err <= tol * scale, (nm, err, scale, (N, C, G, H, W))


def test_fused_train_forward_gradients_match_autocast_module():
    """Whole-network check: loss gradients w.r.t. every parameter through the fused training forward
    vs the unchanged module under fp16 autocast (the reference's ppo_update path, ppo.py:24-26)."""
    import torch
    import minesweeper_ppo_b200 as m
    from minesweeper_ppo_b200.fused_train import fused_train_forward
    torch.manual_seed(3)
    net = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                        model_cfg=dict(stem_channels=96, blocks=3, dropout=0.0, value_hidden=64)).cuda()
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.GroupNorm):
                mod.weight.uniform_(0.5, 1.5); mod.bias.uniform_(-0.3, 0.3)
    x = (torch.rand(48, 10, 16, 16, device="cuda") < 0.3).float()
    wl, wv, wm = torch.randn(48, 256, device="cuda"), torch.randn(48, device="cuda"), torch.randn(48, 1, 16, 16, device="cuda")

    def loss_of(outs):
        l, v, mi = outs
        return (l.float() * wl).mean() + (v.float() * wv).mean() + (mi.float() * wm).mean()

    with torch.autocast("cuda", dtype=torch.float16):
        ref_out = net(x, return_mine=True)
    loss_of(ref_out).backward()
    ref = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    out = fused_train_forward(net, x, return_mine=True)
    for a, b in zip(out, ref_out):
        assert float((a.float() - b.float()).abs().max()) <= 2e-2 * (float(b.float().abs().max()) + 1e-3)
    loss_of(out).backward()
    for k, p in net.named_parameters():
        g, r = p.grad.float().reshape(-1), ref[k].float().reshape(-1)
        cos = float(torch.dot(g, r) / (g.norm() * r.norm() + 1e-30))
        rel = float((g - r).norm() / (r.norm() + 1e-30))
        assert cos > 0.995 and rel < 0.1, (k, cos, rel)      # both paths carry fp16 rounding noise


def test_all_masked_row_guard():
    """train_rl.py:166-168 / 263-265: a row whose mask has no legal action becomes all-legal before sampling
    (and that is the mask the buffer stores).  msw_masked_sample applies the same guard in the launch: the
    row's mask bytes are set, the action is drawn from the raw logits, logp is the unmasked log-softmax."""
    import torch
    from minesweeper_ppo_b200.rollout import masked_sample
    g = torch.Generator(device="cuda").manual_seed(0)
    n, A = 64, 256
    logits = torch.randn((n, A), device="cuda", generator=g)
    mask = torch.rand((n, A), device="cuda", generator=g) < 0.5
    dead = [3, 17, 63]
    mask[dead] = False
    before = mask.clone()
    a64, a32, logp = masked_sample(logits, mask, seed=1, step_index=2)
    assert bool(mask[dead].all()), "guarded rows must come back all-legal"
    keep = [i for i in range(n) if i not in dead]
    assert torch.equal(mask[keep], before[keep]), "other rows untouched"
    want = torch.log_softmax(logits[dead], dim=-1).gather(1, a64[dead].unsqueeze(1)).squeeze(1)
    assert float((logp[dead] - want).abs().max()) < 2e-5
    assert bool(before[keep].gather(1, a64[keep].unsqueeze(1)).all()), "live rows sample legal actions only"
    # the guarded rows sample from ALL cells: over many draws both halves of the row are hit
    hits = torch.zeros(A, device="cuda")
    for s in range(200):
        m2 = torch.zeros((1, A), dtype=torch.bool, device="cuda")
        a, _, _ = masked_sample(torch.zeros((1, A), device="cuda"), m2, seed=s, step_index=0)
        hits[a[0]] += 1
    assert int((hits > 0).sum()) > 100
