"""Development: msw_gae device time (CUDA events around 20 graph-replayed launches, so the host is not in the
way) for several N at T=128.  Run once with MSW_GAE_TMA=0 and once with =1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minesweeper_ppo_b200 as m

dev = torch.device("cuda", 0)
T = 128
w = torch.empty(1 << 28, dtype=torch.uint8, device=dev)
for _ in range(300):
    w.zero_()
for N in (8192, 65536, 524288):
    buf = m.RolloutBuffer(N, T, (1, 1, 1), 1, dev)
    buf.dones.copy_(torch.rand((T * N,), device=dev) < 0.15)
    buf.rewards.copy_(torch.where(buf.dones, torch.tensor(-1.0001, device=dev), torch.tensor(-1e-4, device=dev)))
    buf.values.copy_(0.5 * torch.randn((T * N,), device=dev))
    last = 0.5 * torch.randn((N,), device=dev)
    buf.compute_gae(last, 0.995, 0.95)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        buf.compute_gae(last, 0.995, 0.95)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                buf.compute_gae(last, 0.995, 0.95)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    nbytes = 17 * T * N + 4 * N
    print(f"tma={os.environ.get('MSW_GAE_TMA', 'default')} N={N}: {us:.2f} us per launch (graph of 20, "
          f"{'L2-resident' if nbytes < 100e6 else 'HBM'}), {nbytes / us / 1e3:.0f} GB/s")
