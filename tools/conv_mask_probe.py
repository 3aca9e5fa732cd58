"""Development: per-tap exactness of msw_conv3x3 (one accumulator, row-shifted A descriptors, disable-output-lane
masks), separately for the 128-byte- and the 64-byte-swizzled channel block.  profiles/r01j_conv_descriptor_probe.txt
holds its output for the two descriptor variants tried in round 1."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from minesweeper_ppo_b200.fused_forward import conv3x3, conv3x3_taps

C, n = 96, 4
x = torch.randn((n, C, 16, 16), device="cuda").half().contiguous(memory_format=torch.channels_last)
for lo, hi, name in ((0, 64, "ci 0..63 (128B swizzle)"), (64, 96, "ci 64..95 (64B swizzle)")):
    for ky in range(3):
        row = []
        for kx in range(3):
            w1 = torch.zeros((C, C, 3, 3), device="cuda", dtype=torch.float16)
            w1[:, :, ky, kx][torch.arange(lo, hi), torch.arange(lo, hi)] = 1.0
            got = conv3x3(x, conv3x3_taps(w1))
            ref = F.conv2d(x, w1.contiguous(memory_format=torch.channels_last), None, padding=1)
            bad = (got != ref)
            row.append(f"{int(bad.sum())}")
        print(name, "ky", ky, "mismatches per kx:", " ".join(row))
