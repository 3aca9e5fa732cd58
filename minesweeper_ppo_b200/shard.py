"""Env-id sharding over ranks: envs are independent, so rank g of G owns the
contiguous global-id range [g*N/G, (g+1)*N/G) and the env path needs no collective.  The board
sampler is keyed by GLOBAL env id, so results do not depend on G."""
from __future__ import annotations

from typing import Tuple


def shard_range(num_envs_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(env_id_base, num_local) of `rank`; the first `num_envs_total % world_size` ranks get one more."""
    if not (0 <= rank < world_size) or num_envs_total < world_size:
        raise ValueError(f"bad shard request: total={num_envs_total} rank={rank} world={world_size}")
    q, r = divmod(num_envs_total, world_size)
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def owner_of(env_id: int, num_envs_total: int, world_size: int) -> Tuple[int, int]:
    """(rank, local index) of a global env id under `shard_range`."""
    q, r = divmod(num_envs_total, world_size)
    cut = r * (q + 1)
    if env_id < cut:
        return env_id // (q + 1), env_id % (q + 1)
    return r + (env_id - cut) // q, (env_id - cut) % q
