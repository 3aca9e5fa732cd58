import os, sys, faulthandler
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minesweeper_ppo_b200 as m
from minesweeper_ppo_b200.fused_forward import FusedRolloutForward
torch.manual_seed(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 96
model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                      model_cfg=dict(stem_channels=C, blocks=2, dropout=0.05, value_hidden=64)).cuda()
x = torch.zeros(128, 10, 16, 16, device="cuda")
ff = FusedRolloutForward(model)
if "--default-first" in sys.argv:
    print("default stream forward", flush=True)
    ff(x, return_mine=True); torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
print("side stream forward", flush=True)
with torch.cuda.stream(side):
    ff(x, return_mine=True)
torch.cuda.synchronize()
print("side ok", flush=True)
