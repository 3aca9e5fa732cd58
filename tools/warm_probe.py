"""How long does the env step kernel take per launch as a function of how long the GPU has been running it back to
back?  One CUDA event every 20 launches for 1,200 launches after an idle period (bench.py times 20 launches after a
32-step pre-roll and 5 warm-up steps, and measures 108.6 us per step; over 5,000 steps it measures 103.1 us)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import minesweeper_ppo_b200 as m

N = 65536
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cfg = m.EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
vec = m.VecMinesweeper(N, cfg, seed=0, api="torch")
slots = [m.StepOut(obs=torch.empty((N, 10, 16, 16), dtype=torch.float32, device=dev),
                   action_mask=torch.empty((N, 256), dtype=torch.bool, device=dev),
                   rewards=torch.empty((N,), dtype=torch.float32, device=dev),
                   dones=torch.empty((N,), dtype=torch.bool, device=dev)) for _ in range(4)]
scratch = torch.empty((N,), dtype=torch.int32, device=dev)
vec.reset(out=slots[0])
for rep in range(3):
    torch.cuda.synchronize()
    time.sleep(0.5)                                   # idle GPU
    CH, NCH = 20, 60
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(NCH + 1)]
    t = rep * 10000
    evs[0].record()
    for c in range(NCH):
        for k in range(CH):
            vec.step_random(t, out=slots[t % 4], actions_out=scratch); t += 1
        evs[c + 1].record()
    torch.cuda.synchronize()
    per = [evs[c].elapsed_time(evs[c + 1]) / CH * 1e3 for c in range(NCH)]
    print("rep %d: us per launch by chunk of 20 launches:" % rep, " ".join("%.1f" % x for x in per[:12]), "... mean of chunks 12-59: %.1f" % np.mean(per[12:]))

# finer: one event every 5 launches for the first 60 launches after a synchronize (no sleep) and after a 5 ms sleep
for gap in (0.0, 0.005):
    for rep in range(2):
        torch.cuda.synchronize()
        if gap:
            time.sleep(gap)
        CH, NCH = 5, 12
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(NCH + 1)]
        evs[0].record()
        for c in range(NCH):
            for k in range(CH):
                vec.step_random(t, out=slots[t % 4], actions_out=scratch); t += 1
            evs[c + 1].record()
        torch.cuda.synchronize()
        per = [evs[c].elapsed_time(evs[c + 1]) / CH * 1e3 for c in range(NCH)]
        print("after synchronize + %.0f ms idle: us per launch by chunk of 5 launches:" % (gap * 1e3), " ".join("%.1f" % x for x in per))
