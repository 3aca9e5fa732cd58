"""B200-native rollout hot path of yakvrz/minesweeper-ppo.

Public surface (same names as the reference modules it stands in for):
    EnvConfig, VecMinesweeper      <- minesweeper/env.py
    RolloutBuffer                  <- minesweeper/buffers.py
Everything computes in hand-written sm_100a CUDA kernels behind the C ABI of
include/msw_b200.h (libmsw_b200.so); there is no CPU fallback.
"""
from .env import EnvConfig, StepOut, VecMinesweeper, pack_boards, reward_constants  # noqa: F401
from .buffers import CompactRolloutBuffer, RolloutBuffer  # noqa: F401
from .policy import build_model  # noqa: F401
from .rollout import RolloutCollector, collect_rollout, masked_sample  # noqa: F401
from .shard import shard_range  # noqa: F401

__all__ = ["EnvConfig", "VecMinesweeper", "RolloutBuffer", "StepOut", "pack_boards", "reward_constants",
           "CompactRolloutBuffer", "build_model", "RolloutCollector", "collect_rollout", "masked_sample", "shard_range"]
