"""Development: one fused forward at C3 size (target for ncu on gn_act_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minesweeper_ppo_b200 as m
from minesweeper_ppo_b200.fused_forward import FusedRolloutForward
model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                      model_cfg=dict(stem_channels=96, blocks=5, dropout=0.05, value_hidden=256)).cuda()
x = (torch.rand(8192, 10, 16, 16, device="cuda") < 0.3).float()
ff = FusedRolloutForward(model)
for _ in range(3):
    ff(x, return_mine=True)
torch.cuda.synchronize()
print("ok")
