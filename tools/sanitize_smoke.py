"""Tiny exercise of every kernel in libmsw_b200.so (odd shapes, tails) -- the target of
`compute-sanitizer --tool memcheck`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import minesweeper_ppo_b200 as m
from minesweeper_ppo_b200.fused_forward import FusedRolloutForward

torch.manual_seed(0)
for (H, W, M, N) in [(16, 16, 40, 37), (16, 30, 99, 19), (5, 7, 6, 33), (32, 32, 200, 5), (1, 12, 3, 9)]:
    cfg = m.EnvConfig(H=H, W=W, mine_count=M, step_penalty=1e-4)
    vec = m.VecMinesweeper(N, cfg, seed=1, api="torch", aux_maps=True, late_start_cfg=dict(prob=0.5, min_hidden=1, max_hidden=4))
    vec.reset()
    for t in range(6):
        a = vec.random_actions(t, valid_only=bool(t % 2))
        vec.step(a)
        vec.step_random(t + 100)
    vec._unpacked(); vec.encode()
    comp = m.CompactRolloutBuffer(vec, 2, aux_maps=True)
    comp.snapshot(0); comp.snapshot(1)
    comp.gather_obs(torch.randperm(2 * N, device="cuda")[: N + 3])
    nv = m.VecMinesweeper(N, cfg, seed=1)                      # NumPy API / msw_step_host
    b = nv.reset()
    for t in range(3):                                         # recycled result sets: full write, then delta updates
        b, _, _, _ = nv.step(np.full(N, t, np.int32))
    logits = torch.randn(N, H * W, device="cuda").half()
    m.masked_sample(logits, torch.rand(N, H * W, device="cuda") < 0.5, seed=1, step_index=2)
for T, N in [(1, 1), (129, 33), (5, 70), (70, 48), (161, 16), (40, 16384)]:     # plain-load and both TMA launch shapes
    buf = m.RolloutBuffer(N, T, (1, 1, 1), 1, torch.device("cuda"))
    buf.rewards.normal_(); buf.values.normal_()
    buf.compute_gae(torch.randn(N, device="cuda"))
    buf.compute_gae(torch.randn(N, device="cuda").half())
net = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                    model_cfg=dict(stem_channels=32, blocks=1, dropout=0.1, value_hidden=16)).cuda()
FusedRolloutForward(net)(torch.zeros(3, 10, 16, 16, device="cuda"), return_mine=True)
torch.cuda.synchronize()
print("sanitize smoke ok")
