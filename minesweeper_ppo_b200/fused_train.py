"""Training-time forward of `CNNResidualPolicy` with the fused GroupNorm kernels and autograd
(the training half of SURVEY section 8, row f4).

Same structure and rounding points as `fused_forward.FusedRolloutForward` / the reference's fp16
autocast forward (convolutions and linears in fp16 on the tensor cores, GroupNorm arithmetic and the
residual stream in fp32), but differentiable: `GnAct` is a `torch.autograd.Function` whose forward is
msw_gn_act (saving the group statistics and a ReLU/Dropout2d bit mask) and whose backward is
msw_gn_act_bwd.  cuDNN does the convolution forward / dgrad / wgrad; parameters stay the module's
fp32 tensors (the fp16 casts are part of the graph), so AdamW / GradScaler / clipping are unchanged.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from .fused_forward import conv3x3, conv3x3_taps
from .policy import CNNResidualPolicy

_CL = torch.channels_last


class GnAct(torch.autograd.Function):
    """(y16[, y32]) = relu(GroupNorm(x16 + conv_bias) [+ res32]) [* Dropout2d]  -- see msw_gn.cu."""

    @staticmethod
    def forward(ctx, x16, conv_bias, gamma, beta, res32, groups, eps, relu, drop_p, seed, call_id, want32):
        L = _lib.load()
        x16 = x16.contiguous(memory_format=_CL)
        N, C, H, W = x16.shape
        dev = x16.device
        y16 = torch.empty_like(x16, memory_format=_CL)
        y32 = torch.empty((N, C, H, W), dtype=torch.float32, device=dev, memory_format=_CL) if want32 else None
        mean = torch.empty((N, groups), dtype=torch.float32, device=dev)
        rstd = torch.empty((N, groups), dtype=torch.float32, device=dev)
        mask = torch.empty((N, H * W, C // 8), dtype=torch.uint8, device=dev)
        if res32 is not None:
            res32 = res32.contiguous(memory_format=_CL)
        cb = conv_bias.detach().float().contiguous()
        with torch.cuda.device(dev):
            rc = L.msw_gn_act(x16.data_ptr(), cb.data_ptr(), None if res32 is None else res32.data_ptr(),
                              gamma.data_ptr(), beta.data_ptr(), y16.data_ptr(), None if y32 is None else y32.data_ptr(),
                              N, H * W, C, groups, float(eps), int(relu), float(drop_p), int(seed) & (2**64 - 1),
                              int(call_id) & (2**64 - 1), None, mean.data_ptr(), rstd.data_ptr(), mask.data_ptr(),
                              None, 0, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "msw_gn_act")
        ctx.save_for_backward(x16, cb, gamma, mean, rstd, mask)
        ctx.meta = (groups, float(drop_p), res32 is not None)
        return (y16, y32) if want32 else y16

    @staticmethod
    def backward(ctx, g16, g32=None):
        L = _lib.load()
        x16, cb, gamma, mean, rstd, mask = ctx.saved_tensors
        groups, drop_p, had_res = ctx.meta
        N, C, H, W = x16.shape
        dev = x16.device
        if g16 is not None:
            g16 = g16.to(torch.float16).contiguous(memory_format=_CL)
        if g32 is not None:
            g32 = g32.float().contiguous(memory_format=_CL)
        dx = torch.empty_like(x16, memory_format=_CL)
        dres = torch.empty((N, C, H, W), dtype=torch.float32, device=dev, memory_format=_CL) if had_res else None
        parts = torch.empty((3, N, C), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = L.msw_gn_act_bwd(x16.data_ptr(), cb.data_ptr(), gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                  mask.data_ptr(), None if g16 is None else g16.data_ptr(),
                                  None if g32 is None else g32.data_ptr(), dx.data_ptr(),
                                  None if dres is None else dres.data_ptr(), parts[0].data_ptr(), parts[1].data_ptr(),
                                  parts[2].data_ptr(), N, H * W, C, groups, drop_p,
                                  torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "msw_gn_act_bwd")
        sums = parts.sum(dim=1)                  # fixed-order reduction over the batch
        return dx, sums[2], sums[0], sums[1], dres, None, None, None, None, None, None, None


def gn_act_train(x16, conv: torch.nn.Conv2d, norm: torch.nn.GroupNorm, *, res32=None, drop_p=0.0, seed=0, call_id=0,
                 want32=False):
    return GnAct.apply(x16, conv.bias, norm.weight, norm.bias, res32, norm.num_groups, norm.eps, True, drop_p, seed,
                       call_id, want32)


def supports(model) -> bool:
    if not isinstance(model, CNNResidualPolicy):
        return False
    C, G = model.stem[0].out_channels, model.stem[1].num_groups
    return C % 8 == 0 and (C // G) % 8 == 0


class Conv3x3Tc(torch.autograd.Function):
    """3x3 trunk convolution on the tcgen05 kernel (msw_conv3x3) with its data gradient on the same kernel:
    dL/dx is the convolution of dL/dy with the taps flipped in space and transposed in the channels.  The weight
    gradient stays cuDNN's wgrad (fp16 operands, as the autocast path computes it)."""

    @staticmethod
    def forward(ctx, x16, weight):
        ctx.save_for_backward(x16, weight)
        return conv3x3(x16, conv3x3_taps(weight))

    @staticmethod
    def backward(ctx, gy):
        x16, weight = ctx.saved_tensors
        gy = gy.to(torch.float16).contiguous(memory_format=_CL)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            co, ci = weight.shape[0], weight.shape[1]
            taps_t = weight.detach().flip(2, 3).permute(2, 3, 1, 0).to(torch.float16).reshape(9, ci, co).contiguous()
            gx = conv3x3(gy, taps_t)
        if ctx.needs_input_grad[1]:
            gw = torch.nn.grad.conv2d_weight(x16, weight.shape, gy, padding=1).to(weight.dtype)
        return gx, gw


def _conv3(x16: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    if conv.in_channels == 96 and conv.out_channels == 96 and tuple(x16.shape[2:]) == (16, 16):
        return Conv3x3Tc.apply(x16, conv.weight)
    w16 = conv.weight.to(torch.float16).contiguous(memory_format=_CL)      # differentiable cast
    return F.conv2d(x16, w16, None, padding=1)


def _lin(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return F.linear(x, w.reshape(w.shape[0], -1).to(torch.float16), b.to(torch.float16))


def fused_train_forward(model: CNNResidualPolicy, obs: torch.Tensor, return_mine: bool = False, *, seed: int = 0,
                        step: int = 0):
    """Differentiable forward with the module's `(obs, return_mine)` contract; `step` must change
    between optimizer steps so Dropout2d masks differ (train mode only)."""
    m = model
    x = obs.to(dtype=torch.float16, memory_format=_CL)
    a16, a32 = gn_act_train(_conv3(x, m.stem[0]), m.stem[0], m.stem[1], want32=True)
    for k, blk in enumerate(m.residual_stack):
        p = float(blk.dropout.p) if (m.training and isinstance(blk.dropout, torch.nn.Dropout2d)) else 0.0
        t16 = gn_act_train(_conv3(a16, blk.conv1), blk.conv1, blk.norm1, drop_p=p, seed=seed, call_id=(step << 8) + k)
        a16, a32 = gn_act_train(_conv3(t16, blk.conv2), blk.conv2, blk.norm2, res32=a32, want32=True)
    n, c, h, w = a16.shape
    rows = a16.permute(0, 2, 3, 1).reshape(n * h * w, c)
    ph = m.policy_head
    logits = _lin(F.relu(_lin(rows, ph[0].weight, ph[0].bias)), ph[2].weight, ph[2].bias).reshape(n, h * w)
    v = a32.mean(dim=(2, 3)).to(torch.float16)
    vh = m.value_head
    v = F.relu(_lin(v, vh[2].weight, vh[2].bias))
    v = F.relu(_lin(v, vh[4].weight, vh[4].bias))
    value = _lin(v, vh[6].weight, vh[6].bias).squeeze(-1)
    if not return_mine:
        return logits, value
    mh = m.mine_head
    d = rows.detach()                          # belief head does not train the trunk (cnn_residual.py:94)
    mine = _lin(F.relu(_lin(d, mh[0].weight, mh[0].bias)), mh[2].weight, mh[2].bias).reshape(n, 1, h, w)
    return logits, value, mine
