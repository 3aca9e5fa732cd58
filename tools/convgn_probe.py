"""Development: msw_conv3x3_gn per epilogue mode at C3 size (8,192 boards): no residual (+Dropout2d), residual in +
y32 out (P8 order), residual in + pooled output; against the unfused pair msw_conv3x3 -> msw_gn_act.  Inputs
alternate between two sets so they come from HBM (every operand set > L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from minesweeper_ppo_b200.fused_forward import conv3x3, conv3x3_gn, conv3x3_taps, gn_act, to_p8

n, C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 96
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
xs = [torch.randn((n, C, 16, 16), device="cuda").half().contiguous(memory_format=torch.channels_last) for _ in range(2)]
rs = [torch.randn((n, 2, 3, 4, 128, 8), device="cuda") for _ in range(2)]
rn = [torch.randn((n, C, 16, 16), device="cuda").contiguous(memory_format=torch.channels_last) for _ in range(2)]
w = (torch.randn((C, C, 3, 3), device="cuda") / (9 * C) ** 0.5).half()
taps = conv3x3_taps(w)
bias = 0.1 * torch.randn((C,), device="cuda")
norm = torch.nn.GroupNorm(6, C).cuda()


def timed(fn):
    for i in range(4):
        fn(i & 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i & 1)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


act = n * 256 * C
flops = 2.0 * n * 256 * C * C * 9
cases = [
    ("conv_gn  no residual, dropout 0.05        ", lambda i: conv3x3_gn(xs[i], taps, norm, bias, drop_p=0.05, seed=1, call_id=i), 4 * act),
    ("conv_gn  residual in, y32 out (P8)        ", lambda i: conv3x3_gn(xs[i], taps, norm, bias, res32=rs[i], want32=True), 12 * act),
    ("conv_gn  residual in, pooled out          ", lambda i: conv3x3_gn(xs[i], taps, norm, bias, res32=rs[i], want_pool=True), 8 * act),
    ("unfused  conv3x3 -> gn_act (residual, y32)", lambda i: gn_act(conv3x3(xs[i], taps), norm, conv_bias=bias, res32=rn[i], want32=True), 16 * act),
    ("plain    conv3x3                          ", lambda i: conv3x3(xs[i], taps), 4 * act),
]
for name, fn, nbytes in cases:
    us = timed(fn)
    print(f"{name}: {us:7.1f} us  {flops / us / 1e6:6.0f} TFLOP/s  {nbytes / us / 1e3:6.0f} GB/s algorithmic HBM traffic")
