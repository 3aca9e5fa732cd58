"""Development: run msw_gae at C3 size a few times (cold L2) -- target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
import bench
import minesweeper_ppo_b200 as m
_w = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(300):          # bring the clocks up before timing a 10 us kernel
    _w.zero_()
torch.cuda.synchronize()
print(json.dumps(bench.bench_gae(torch, m, torch.device("cuda", 0))))
