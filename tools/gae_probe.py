"""Development: run msw_gae at C3 size a few times (cold L2) -- target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json
import torch
import bench
import minesweeper_ppo_b200 as m
print(json.dumps(bench.bench_gae(torch, m, torch.device("cuda", 0))))
