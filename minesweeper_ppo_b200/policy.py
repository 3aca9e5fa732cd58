"""Policy / value / mine-belief networks used on the rollout path.

These stay plain PyTorch modules executed by cuDNN on the tensor cores ("kept
PyTorch, not re-implemented"); they exist here only so rollouts can run where the reference tree
is absent.  Architectures and parameter names follow `minesweeper/models/cnn_residual.py:7-96`
and `minesweeper/models/cnn.py:7-60`, so `state_dict`s are interchangeable with reference
checkpoints (`train_rl.py:623-630`); `build_model` has the signature of
`minesweeper/models/__init__.py:17-49`.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn


def _conv3(cin: int, cout: int) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, kernel_size=3, padding=1)


def _pixel_head(ch: int) -> nn.Sequential:
    """1x1 -> ReLU -> 1x1 per-cell head (policy logits / mine logits)."""
    return nn.Sequential(nn.Conv2d(ch, ch, kernel_size=1), nn.ReLU(inplace=True), nn.Conv2d(ch, 1, kernel_size=1))


class _ResidualBlock(nn.Module):
    """conv-GN-ReLU-Dropout2d-conv-GN, skip add, ReLU (cnn_residual.py:7-27)."""

    def __init__(self, channels: int, groups: int, dropout: float = 0.0) -> None:
        super().__init__()
        self.conv1 = _conv3(channels, channels)
        self.norm1 = nn.GroupNorm(groups, channels)
        self.conv2 = _conv3(channels, channels)
        self.norm2 = nn.GroupNorm(groups, channels)
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else nn.Identity()
        self.act = nn.ReLU(inplace=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.dropout(self.act(self.norm1(self.conv1(x))))
        y = self.norm2(self.conv2(y))
        return self.act(y + x)


class CNNResidualPolicy(nn.Module):
    """Residual trunk with policy, value (GAP -> MLP) and detached mine heads (cnn_residual.py:30-96)."""

    def __init__(self, in_channels: int, *, stem_channels: int = 128, blocks: int = 6, dropout: float = 0.05,
                 value_hidden: int = 256) -> None:
        super().__init__()
        if stem_channels <= 0 or blocks <= 0:
            raise ValueError("stem_channels and blocks must be positive")
        groups = max(1, stem_channels // 16)
        self.stem = nn.Sequential(_conv3(in_channels, stem_channels), nn.GroupNorm(groups, stem_channels),
                                  nn.ReLU(inplace=True))
        self.residual_stack = nn.Sequential(*[_ResidualBlock(stem_channels, groups, dropout) for _ in range(blocks)])
        self.policy_head = _pixel_head(stem_channels)
        self.value_head = nn.Sequential(
            nn.AdaptiveAvgPool2d(1), nn.Flatten(),
            nn.Linear(stem_channels, value_hidden), nn.ReLU(inplace=True),
            nn.Linear(value_hidden, value_hidden), nn.ReLU(inplace=True),
            nn.Linear(value_hidden, 1),
        )
        self.mine_head = _pixel_head(stem_channels)

    def set_gradient_checkpointing(self, enabled: bool) -> None:
        return None

    def forward(self, x: torch.Tensor, return_mine: bool = False):
        feat = self.residual_stack(self.stem(x))
        n, _, h, w = feat.shape
        logits = self.policy_head(feat).permute(0, 2, 3, 1).reshape(n, h * w)     # row-major cell order
        value = self.value_head(feat).squeeze(-1)
        if not return_mine:
            return logits, value
        return logits, value, self.mine_head(feat.detach())      # belief head does not train the trunk

    def beta_regularizer(self) -> torch.Tensor:
        return next(self.parameters()).new_zeros(())


class CNNPolicy(nn.Module):
    """Three-conv baseline (cnn.py:7-60)."""

    def __init__(self, in_channels: int, hidden: int = 64) -> None:
        super().__init__()
        hidden = int(hidden)
        if hidden <= 0:
            raise ValueError("hidden must be positive")
        width = 64
        self.backbone = nn.Sequential(
            _conv3(in_channels, 32), nn.ReLU(inplace=True), nn.GroupNorm(4, 32),
            _conv3(32, 64), nn.ReLU(inplace=True), nn.GroupNorm(8, 64),
            _conv3(64, width), nn.ReLU(inplace=True),
        )
        self.policy_head = nn.Conv2d(width, 1, kernel_size=1)
        self.value_head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(width, hidden),
                                        nn.ReLU(inplace=True), nn.Linear(hidden, 1))
        self.mine_head = nn.Conv2d(width, 1, kernel_size=1)

    def set_gradient_checkpointing(self, enabled: bool) -> None:
        return None

    def forward(self, x: torch.Tensor, return_mine: bool = False):
        feat = self.backbone(x)
        n, _, h, w = feat.shape
        logits = self.policy_head(feat).permute(0, 2, 3, 1).reshape(n, h * w)
        value = self.value_head(feat).squeeze(-1)
        if not return_mine:
            return logits, value
        return logits, value, self.mine_head(feat)

    def beta_regularizer(self) -> torch.Tensor:
        return next(self.parameters()).new_zeros(())


def build_model(name: str, *, obs_shape: Tuple[int, int, int], env_overrides: Optional[Dict[str, bool]] = None,
                model_cfg: Optional[dict] = None) -> nn.Module:
    """Same contract as models/__init__.py:17-49 (defaults included)."""
    cfg = dict(model_cfg or {})
    cin = obs_shape[0]
    if name == "cnn":
        return CNNPolicy(in_channels=cin, hidden=int(cfg.pop("hidden", 64)))
    if name in ("cnn_residual", "cnn_large"):
        return CNNResidualPolicy(
            in_channels=cin,
            stem_channels=int(cfg.pop("stem_channels", 128)),
            blocks=int(cfg.pop("blocks", 6)),
            dropout=float(cfg.pop("dropout", 0.05)),
            value_hidden=int(cfg.pop("value_hidden", 256)),
        )
    raise ValueError(f"Unknown model name: {name}")
