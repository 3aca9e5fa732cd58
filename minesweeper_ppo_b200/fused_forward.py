"""Inference-only forward of `CNNResidualPolicy` for rollouts: cuDNN fp16 NHWC convolutions with
ONE fused kernel (msw_gn_act) between them instead of the eager GroupNorm / ReLU / Dropout2d /
residual-add / dtype-cast kernels.

Numerics follow the reference's fp16-autocast forward (train_rl.py:222-227 over
cnn_residual.py:83-96): convolutions and linears in fp16, GroupNorm statistics and arithmetic in
fp32, residual stream in fp32, activations rounded to fp16 exactly where autocast rounds them (at
the next convolution's input).  Dropout2d uses the package's counter-based RNG, not torch's stream.
The parameters are read from the live module, so training and rollouts share one set of weights.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .policy import CNNResidualPolicy


def gn_act(x16: torch.Tensor, norm: torch.nn.GroupNorm, *, res32: Optional[torch.Tensor] = None, relu: bool = True,
           drop_p: float = 0.0, want16: bool = True, want32: bool = False, seed: int = 0, call_id: int = 0
           ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """x16: fp16 [N,C,H,W] in channels_last memory format (conv output).  Returns (y16, y32) with the
    same logical shape / memory format."""
    L = _lib.load()
    if x16.dtype != torch.float16 or not x16.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("gn_act: x must be fp16 channels_last")
    N, C, H, W = x16.shape
    dev = x16.device
    y16 = torch.empty_like(x16, memory_format=torch.channels_last) if want16 else None
    y32 = (torch.empty((N, C, H, W), dtype=torch.float32, device=dev).contiguous(memory_format=torch.channels_last)
           if want32 else None)
    if res32 is not None and (res32.dtype != torch.float32 or tuple(res32.shape) != (N, C, H, W)
                              or not res32.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("gn_act: residual must be fp32 channels_last of the same shape")
    with torch.cuda.device(dev):
        rc = L.msw_gn_act(x16.data_ptr(), None if res32 is None else res32.data_ptr(), norm.weight.data_ptr(),
                          norm.bias.data_ptr(), None if y16 is None else y16.data_ptr(),
                          None if y32 is None else y32.data_ptr(), N, H * W, C, norm.num_groups, float(norm.eps),
                          int(relu), float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(call_id) & 0xFFFFFFFFFFFFFFFF,
                          torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "msw_gn_act")
    return y16, y32


class FusedRolloutForward:
    """Callable with the module's `(obs, return_mine)` signature, for use under torch.no_grad()."""

    def __init__(self, model: CNNResidualPolicy, seed: int = 0):
        if not isinstance(model, CNNResidualPolicy):
            raise TypeError("FusedRolloutForward supports CNNResidualPolicy only")
        C = model.stem[0].out_channels
        G = model.stem[1].num_groups
        if C % 8 or (C // G) % 8:
            raise ValueError(f"fused forward needs C % 8 == 0 and (C/G) % 8 == 0, got C={C} G={G}")
        self.model, self.seed, self.calls = model, int(seed), 0
        self._w: List[torch.Tensor] = []
        self.refresh()

    @staticmethod
    def supports(model) -> bool:
        if not isinstance(model, CNNResidualPolicy):
            return False
        C, G = model.stem[0].out_channels, model.stem[1].num_groups
        return C % 8 == 0 and (C // G) % 8 == 0

    @torch.no_grad()
    def refresh(self) -> None:
        """Re-cast the (possibly just updated) fp32 parameters to the fp16 copies the convs use."""
        m = self.model

        def conv(c):
            return (c.weight.detach().to(torch.float16).contiguous(memory_format=torch.channels_last),
                    None if c.bias is None else c.bias.detach().to(torch.float16))

        def lin(l):
            return l.weight.detach().to(torch.float16), l.bias.detach().to(torch.float16)

        self.stem = conv(m.stem[0])
        self.blocks = [(conv(b.conv1), conv(b.conv2)) for b in m.residual_stack]
        self.policy = (conv(m.policy_head[0]), conv(m.policy_head[2]))
        self.mine = (conv(m.mine_head[0]), conv(m.mine_head[2]))
        self.value = [lin(m.value_head[i]) for i in (2, 4, 6)]

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, return_mine: bool = False):
        m = self.model
        self.calls += 1
        cid = self.calls << 8
        x = obs.to(dtype=torch.float16, memory_format=torch.channels_last)
        y = F.conv2d(x, *self.stem, padding=1)
        a16, a32 = gn_act(y, m.stem[1], relu=True, want32=True)
        for k, (blk, (c1, c2)) in enumerate(zip(m.residual_stack, self.blocks)):
            p = float(blk.dropout.p) if (m.training and isinstance(blk.dropout, torch.nn.Dropout2d)) else 0.0
            t16, _ = gn_act(F.conv2d(a16, *c1, padding=1), blk.norm1, relu=True, drop_p=p, seed=self.seed,
                            call_id=cid + k)
            a16, a32 = gn_act(F.conv2d(t16, *c2, padding=1), blk.norm2, res32=a32, relu=True, want32=True)
        n, _, h, w = a16.shape
        logits = F.conv2d(F.relu_(F.conv2d(a16, *self.policy[0])), *self.policy[1])
        logits = logits.permute(0, 2, 3, 1).reshape(n, h * w)
        pooled = a32.mean(dim=(2, 3))                              # AdaptiveAvgPool2d(1) in fp32
        v = pooled.to(torch.float16)
        v = F.relu_(F.linear(v, *self.value[0]))
        v = F.relu_(F.linear(v, *self.value[1]))
        value = F.linear(v, *self.value[2]).squeeze(-1)
        if not return_mine:
            return logits, value
        mine = F.conv2d(F.relu_(F.conv2d(a16, *self.mine[0])), *self.mine[1])
        return logits, value, mine
