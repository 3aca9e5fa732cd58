"""Development: msw_conv3x3 (tcgen05) vs cuDNN for the trunk convolution at C3 size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from minesweeper_ppo_b200.fused_forward import conv3x3, conv3x3_taps

n, C = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 96
x = torch.randn((n, C, 16, 16), device="cuda").half().contiguous(memory_format=torch.channels_last)
w = (torch.randn((C, C, 3, 3), device="cuda") / (9 * C) ** 0.5).half()
wcl, taps = w.contiguous(memory_format=torch.channels_last), conv3x3_taps(w)
x2 = torch.randn_like(x).contiguous(memory_format=torch.channels_last)   # alternate inputs: 2 x 403 MB > L2


def timed(fn, reps=20):
    for _ in range(3):
        fn(x); fn(x2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(x if i & 1 else x2)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


flops = 2.0 * n * 256 * C * C * 9
for name, fn in (("cuDNN F.conv2d", lambda t: F.conv2d(t, wcl, None, padding=1)), ("msw_conv3x3", lambda t: conv3x3(t, taps))):
    us = timed(fn)
    print(f"{name}: {us:.1f} us  {flops / us / 1e6:.0f} TFLOP/s  {2 * n * 256 * C * 2 / us / 1e3:.0f} GB/s of activation traffic")
