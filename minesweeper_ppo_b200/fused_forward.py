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

import os
from typing import List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .policy import CNNResidualPolicy


def gn_act(x16: torch.Tensor, norm: torch.nn.GroupNorm, *, conv_bias: Optional[torch.Tensor] = None,
           res32: Optional[torch.Tensor] = None, relu: bool = True,
           drop_p: float = 0.0, want16: bool = True, want32: bool = False, seed: int = 0, call_id: int = 0,
           epoch: Optional[torch.Tensor] = None, pool32: Optional[torch.Tensor] = None
           ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """x16: fp16 [N,C,H,W] in channels_last memory format (conv output, bias NOT applied when
    `conv_bias` -- fp32 [C] -- is given).  Returns (y16, y32) with the same logical shape / memory format;
    `pool32` (fp32 [N,C], optional) receives the spatial mean of the fp32 output."""
    L = _lib.load()
    if x16.dtype != torch.float16 or not x16.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("gn_act: x must be fp16 channels_last")
    N, C, H, W = x16.shape
    dev = x16.device
    y16 = torch.empty_like(x16, memory_format=torch.channels_last) if want16 else None
    y32 = (torch.empty((N, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
           if want32 else None)
    if pool32 is not None and (pool32.dtype != torch.float32 or tuple(pool32.shape) != (N, C) or not pool32.is_contiguous()):
        raise ValueError("gn_act: pool32 must be contiguous fp32 [N, C]")
    if res32 is not None and (res32.dtype != torch.float32 or tuple(res32.shape) != (N, C, H, W)
                              or not res32.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("gn_act: residual must be fp32 channels_last of the same shape")
    with torch.cuda.device(dev):
        rc = L.msw_gn_act(x16.data_ptr(), None if conv_bias is None else conv_bias.data_ptr(),
                          None if res32 is None else res32.data_ptr(), norm.weight.data_ptr(),
                          norm.bias.data_ptr(), None if y16 is None else y16.data_ptr(),
                          None if y32 is None else y32.data_ptr(), N, H * W, C, norm.num_groups, float(norm.eps),
                          int(relu), float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, int(call_id) & 0xFFFFFFFFFFFFFFFF,
                          None if epoch is None else epoch.data_ptr(), None, None, None,
                          None if pool32 is None else pool32.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "msw_gn_act")
    return y16, y32


def cell_heads(rows16: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """msw_cell_heads: rows16 fp16 [R, C] (NHWC activation viewed as rows) -> (policy logit, mine logit),
    fp16 [R] each; w1 [2C, C], b1 [2C], w2 [2C], b2 [2] fp16."""
    L = _lib.load()
    R, C = rows16.shape
    if rows16.dtype != torch.float16 or not rows16.is_contiguous():
        raise ValueError("cell_heads: rows must be contiguous fp16 [R, C]")
    dev = rows16.device
    out = torch.empty((2, R), dtype=torch.float16, device=dev)
    with torch.cuda.device(dev):
        rc = L.msw_cell_heads(rows16.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                              out[0].data_ptr(), out[1].data_ptr(), R, C, torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "msw_cell_heads")
    return out[0], out[1]


def conv3x3_taps(weight: torch.Tensor) -> torch.Tensor:
    """Conv2d weight [C_out, C_in, 3, 3] -> the fp16 [9, C_out, C_in] tap-major layout msw_conv3x3 reads."""
    co, ci = weight.shape[0], weight.shape[1]
    return weight.detach().permute(2, 3, 0, 1).to(torch.float16).reshape(9, co, ci).contiguous()


def conv3x3(x16: torch.Tensor, taps16: torch.Tensor) -> torch.Tensor:
    """msw_conv3x3: x16 fp16 [N,C,16,16] channels_last, taps16 from `conv3x3_taps` -> fp16 conv output
    (no bias), same shape / memory format."""
    L = _lib.load()
    N, Cin, H, W = x16.shape
    if x16.dtype != torch.float16 or not x16.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("conv3x3: x must be fp16 channels_last")
    if taps16.dtype != torch.float16 or taps16.dim() != 3 or taps16.shape[0] != 9 or taps16.shape[2] != Cin or not taps16.is_contiguous():
        raise ValueError("conv3x3: taps must be contiguous fp16 [9, C_out, C_in]")
    C = taps16.shape[1]
    y = torch.empty((N, C, H, W), dtype=torch.float16, device=x16.device, memory_format=torch.channels_last)
    with torch.cuda.device(x16.device):
        rc = L.msw_conv3x3(x16.data_ptr(), taps16.data_ptr(), y.data_ptr(), N, H, W, Cin, C,
                           torch.cuda.current_stream(x16.device).cuda_stream)
    _lib.check(rc, "msw_conv3x3")
    return y


def conv3x3_gn(x16: torch.Tensor, taps16: torch.Tensor, norm: torch.nn.GroupNorm, conv_bias: torch.Tensor, *,
               res32: Optional[torch.Tensor] = None, drop_p: float = 0.0, want32: bool = False, seed: int = 0,
               call_id: int = 0, epoch: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """msw_conv3x3_gn = `gn_act(conv3x3(x16, taps16), norm, conv_bias=..., ...)` in one launch."""
    L = _lib.load()
    N, Cin, H, W = x16.shape
    if x16.dtype != torch.float16 or not x16.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("conv3x3_gn: x must be fp16 channels_last")
    if taps16.dtype != torch.float16 or taps16.dim() != 3 or taps16.shape[0] != 9 or taps16.shape[2] != Cin or not taps16.is_contiguous():
        raise ValueError("conv3x3_gn: taps must be contiguous fp16 [9, C_out, C_in]")
    C = taps16.shape[1]
    if res32 is not None and (res32.dtype != torch.float32 or tuple(res32.shape) != (N, C, H, W)
                              or not res32.is_contiguous(memory_format=torch.channels_last)):
        raise ValueError("conv3x3_gn: residual must be fp32 channels_last of the same shape")
    dev = x16.device
    y16 = torch.empty((N, C, H, W), dtype=torch.float16, device=dev, memory_format=torch.channels_last)
    y32 = (torch.empty((N, C, H, W), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
           if want32 else None)
    with torch.cuda.device(dev):
        rc = L.msw_conv3x3_gn(x16.data_ptr(), taps16.data_ptr(), conv_bias.data_ptr(),
                              None if res32 is None else res32.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(),
                              y16.data_ptr(), None if y32 is None else y32.data_ptr(), N, H, W, Cin, C, norm.num_groups,
                              float(norm.eps), float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF,
                              int(call_id) & 0xFFFFFFFFFFFFFFFF, None if epoch is None else epoch.data_ptr(),
                              torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "msw_conv3x3_gn")
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
        self.epoch: Optional[torch.Tensor] = None     # device uint32 counter mixed into the dropout RNG (graph replays)
        self._w: List = []
        self.refresh()

    @staticmethod
    def supports(model) -> bool:
        if not isinstance(model, CNNResidualPolicy):
            return False
        C, G = model.stem[0].out_channels, model.stem[1].num_groups
        return C % 8 == 0 and (C // G) % 8 == 0

    @torch.no_grad()
    def refresh(self) -> None:
        """Re-cast the (possibly just updated) fp32 parameters into the fp16 copies the convs use.
        The copies are allocated once and updated IN PLACE, so their addresses are stable and the
        whole forward -- including this refresh -- can live inside a captured CUDA graph."""
        m = self.model
        p0, p2, q0, q2 = m.policy_head[0], m.policy_head[2], m.mine_head[0], m.mine_head[2]
        C = p0.out_channels
        if not self._w:
            def conv3(c):    # 3x3 conv: fp16 NHWC weight for cuDNN; the bias is folded into msw_gn_act
                return (torch.empty_like(c.weight, dtype=torch.float16, memory_format=torch.channels_last),
                        torch.empty_like(c.bias, dtype=torch.float32))

            def lin(out_f, in_f, dev):
                return (torch.empty((out_f, in_f), dtype=torch.float16, device=dev),
                        torch.empty((out_f,), dtype=torch.float16, device=dev))

            dev = p0.weight.device
            self.stem = conv3(m.stem[0])
            cin = m.stem[0].in_channels
            # stem weight with the input channels padded to 16 (zeros): pairs with msw_pack_obs16, so cuDNN gets a
            # tensor-core friendly input and runs neither its padding kernels nor a separate cast
            self.stem16 = (torch.zeros((C, 16, 3, 3), dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last)
                           if cin <= 16 else None)
            self.stem_taps = (torch.zeros((9, C, 16), dtype=torch.float16, device=dev)
                              if cin <= 16 and C == 96 and os.environ.get("MSW_CONV", "tc") != "cudnn" else None)
            self.blocks = [(conv3(b.conv1), conv3(b.conv2)) for b in m.residual_stack]
            # tap-major fp16 copies of the trunk weights for msw_conv3x3 (tcgen05; 16x16 boards, 96 channels)
            self.taps = ([(torch.empty((9, C, C), dtype=torch.float16, device=dev),
                           torch.empty((9, C, C), dtype=torch.float16, device=dev)) for _ in m.residual_stack]
                         if C == 96 and os.environ.get("MSW_CONV", "tc") != "cudnn" else None)
            # the two per-cell heads (1x1 -> ReLU -> 1x1) are row-wise linears on the NHWC activation:
            # one msw_cell_heads launch, the hidden layer never reaches HBM
            self.head1 = lin(2 * C, C, dev)
            self.head2 = (torch.empty((2 * C,), dtype=torch.float16, device=dev),       # policy w2 | mine w2
                          torch.empty((2,), dtype=torch.float16, device=dev))
            self.value = [lin(m.value_head[i].out_features, m.value_head[i].in_features, dev) for i in (2, 4, 6)]
            self._w = [self.stem]

        def put3(dst, c):
            dst[0].copy_(c.weight)
            dst[1].copy_(c.bias)

        put3(self.stem, m.stem[0])
        if self.stem16 is not None:
            self.stem16[:, :m.stem[0].in_channels].copy_(m.stem[0].weight)
        if self.stem_taps is not None:
            self.stem_taps[:, :, :m.stem[0].in_channels].copy_(m.stem[0].weight.permute(2, 3, 0, 1).reshape(9, C, -1))
        for (d1, d2), b in zip(self.blocks, m.residual_stack):
            put3(d1, b.conv1)
            put3(d2, b.conv2)
        if self.taps is not None:
            for (t1, t2), b in zip(self.taps, m.residual_stack):
                t1.copy_(b.conv1.weight.permute(2, 3, 0, 1).reshape(9, C, C))
                t2.copy_(b.conv2.weight.permute(2, 3, 0, 1).reshape(9, C, C))
        self.head1[0][:C].copy_(p0.weight.reshape(C, C)); self.head1[0][C:].copy_(q0.weight.reshape(C, C))
        self.head1[1][:C].copy_(p0.bias); self.head1[1][C:].copy_(q0.bias)
        self.head2[0][:C].copy_(p2.weight.reshape(-1)); self.head2[0][C:].copy_(q2.weight.reshape(-1))
        self.head2[1][0:1].copy_(p2.bias); self.head2[1][1:2].copy_(q2.bias)
        for dst, i in zip(self.value, (2, 4, 6)):
            dst[0].copy_(m.value_head[i].weight)
            dst[1].copy_(m.value_head[i].bias)

    @torch.no_grad()
    def __call__(self, obs: torch.Tensor, return_mine: bool = False):
        m = self.model
        self.calls += 1
        cid = self.calls << 8
        if self.stem16 is not None and obs.dtype == torch.float32 and obs.is_contiguous():
            nb, cin, hh, ww = obs.shape
            x = torch.empty((nb, 16, hh, ww), dtype=torch.float16, device=obs.device, memory_format=torch.channels_last)
            with torch.cuda.device(obs.device):
                _lib.check(_lib.load().msw_pack_obs16(obs.data_ptr(), x.data_ptr(), nb, cin, hh * ww,
                                                      torch.cuda.current_stream(obs.device).cuda_stream), "msw_pack_obs16")
            if (self.stem_taps is not None and (hh, ww) == (16, 16) and m.stem[1].num_groups == 6
                    and os.environ.get("MSW_CONV_GN", "1") != "0"):
                c0 = None                        # stem conv + GroupNorm + ReLU in one launch
                a16, a32 = conv3x3_gn(x, self.stem_taps, m.stem[1], self.stem[1], want32=True)
            else:
                c0 = F.conv2d(x, self.stem16, None, padding=1)
        else:
            c0 = F.conv2d(obs.to(dtype=torch.float16, memory_format=torch.channels_last), self.stem[0], None, padding=1)
        if c0 is not None:
            a16, a32 = gn_act(c0, m.stem[1], conv_bias=self.stem[1], want32=True)
        last = len(self.blocks) - 1
        pooled = None
        own_conv = self.taps is not None and tuple(a16.shape[2:]) == (16, 16)
        # msw_conv3x3_gn (GroupNorm fused into the conv epilogue).  Measured per layer at 8,192 boards: without a
        # residual 335 us fused vs 239 + 165 us; with the residual stream ~760 us fused vs 239 + 421 us (the
        # thread-per-pixel epilogue reads / writes the fp32 stream uncoalesced).  MSW_CONV_GN: 1 (default) fuses
        # the first half of every block, 2 both halves, 0 none.
        gn_mode = int(os.environ.get("MSW_CONV_GN", "1")) if own_conv and m.stem[1].num_groups == 6 else 0
        for k, (blk, ((w1, b1), (w2, b2))) in enumerate(zip(m.residual_stack, self.blocks)):
            p = float(blk.dropout.p) if (m.training and isinstance(blk.dropout, torch.nn.Dropout2d)) else 0.0
            if gn_mode >= 1:             # conv + GroupNorm + ReLU + Dropout2d in one launch
                t16, _ = conv3x3_gn(a16, self.taps[k][0], blk.norm1, b1, drop_p=p, seed=self.seed, call_id=cid + k,
                                    epoch=self.epoch)
            else:
                c1 = conv3x3(a16, self.taps[k][0]) if own_conv else F.conv2d(a16, w1, None, padding=1)
                t16, _ = gn_act(c1, blk.norm1, conv_bias=b1, drop_p=p, seed=self.seed, call_id=cid + k, epoch=self.epoch)
            if k == last:
                # nothing reads the fp32 residual stream after the last block except the value head's
                # AdaptiveAvgPool2d(1): msw_gn_act emits that mean instead of writing y32 (unfused call)
                pooled = torch.empty((a16.shape[0], a16.shape[1]), dtype=torch.float32, device=a16.device)
            if gn_mode >= 2 and k != last:   # conv + GroupNorm + residual add + ReLU in one launch
                a16, a32 = conv3x3_gn(t16, self.taps[k][1], blk.norm2, b2, res32=a32, want32=True)
            else:
                c2 = conv3x3(t16, self.taps[k][1]) if own_conv else F.conv2d(t16, w2, None, padding=1)
                a16, a32 = gn_act(c2, blk.norm2, conv_bias=b2, res32=a32, want32=k != last, pool32=pooled)
        n, c, h, w = a16.shape
        rows = a16.permute(0, 2, 3, 1).reshape(n * h * w, c)         # NHWC storage: a view, no copy
        if pooled is None:                                           # no residual blocks
            pooled = a32.mean(dim=(2, 3))                            # AdaptiveAvgPool2d(1) in fp32
        if c in (32, 64, 96, 128):
            pol, mine = cell_heads(rows, self.head1[0], self.head1[1], self.head2[0], self.head2[1])
        else:                                                        # other widths: the same math as library GEMMs
            hid = F.relu_(F.linear(rows, *self.head1))               # [n*h*w, 2C]
            pol = F.linear(hid[:, :c], self.head2[0][None, :c], self.head2[1][0:1]).squeeze(-1)
            mine = F.linear(hid[:, c:], self.head2[0][None, c:], self.head2[1][1:2]).squeeze(-1)
        logits = pol.reshape(n, h * w)
        v = F.relu_(F.linear(pooled.to(torch.float16), *self.value[0]))
        v = F.relu_(F.linear(v, *self.value[1]))
        value = F.linear(v, *self.value[2]).squeeze(-1)
        if not return_mine:
            return logits, value
        return logits, value, mine.reshape(n, 1, h, w)
