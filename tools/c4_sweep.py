#!/usr/bin/env python
"""BASELINE.json configs[3] (C4): Expert boards (H=16, W=30, 99 mines), env-only random valid actions, sweep
of the TOTAL env count N over the GPUs of this launch (N/G envs per GPU, env shards, no data-path collective).

    python tools/c4_sweep.py                                  # 1 GPU: N = 65,536 ... 4,194,304
    torchrun --nproc-per-node G tools/c4_sweep.py             # G GPUs: the same N, N/G per GPU

Prints one JSON line per N (rank 0): env-steps/s (max-over-ranks device time), per-GPU HBM GB/s at 19,689
algorithmic bytes per env-step and the fraction of the measured HBM peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench
import minesweeper_ppo_b200 as m

rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
world = max(world, 1)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def reduce_max(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


H, W, M = 16, 30, 99
bps = 41 * H * W + 9
peak, _ = bench.measured_peak_gbs()
totals = [int(x) for x in sys.argv[1:]] or [65536 << k for k in range(7)]      # 65,536 ... 4,194,304
for N in totals:
    n = N // world
    if n < 1:
        continue
    ring = 2 if n * bps * 2 < 120e9 else 1
    K, Wm = 12, 3
    ms_total, ms_kernel, _, _ = bench.time_env_steps(torch, m, dev, rank, world, barrier, (H, W, M), n, K, Wm, ring)
    torch.cuda.empty_cache()
    ms = reduce_max(ms_total)
    mk = reduce_max(ms_kernel)
    if rank == 0:
        gbs = bps * n / (mk / 1e3) / 1e9
        print(json.dumps({"workload": "C4", "board": f"{H}x{W}x{M}", "N_total": N, "gpus": world, "envs_per_gpu": n,
                          "env_steps_per_s": n * world * K / (ms / 1e3), "ms_per_step": ms / K, "kernel_ms": mk,
                          "per_gpu_achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak, "ring_slots": ring}), flush=True)
if world > 1:
    dist.destroy_process_group()
