"""Development: where does the (unchanged PyTorch/cuDNN) policy forward spend its time at C3 size?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minesweeper_ppo_b200 as m
from torch.profiler import profile, ProfilerActivity

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda")
model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                      model_cfg=dict(stem_channels=96, blocks=5, dropout=0.05, value_hidden=256)).to(dev)
x = (torch.rand(N, 10, 16, 16, device=dev) < 0.3).float()

def run(tag, model, x, bench=False):
    torch.backends.cudnn.benchmark = bench
    with torch.no_grad():
        for _ in range(3):
            with torch.autocast("cuda", dtype=torch.float16):
                model(x, return_mine=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            with torch.autocast("cuda", dtype=torch.float16):
                model(x, return_mine=True)
        b.record(); torch.cuda.synchronize()
    print(f"{tag}: {a.elapsed_time(b)/5:.2f} ms / forward")

run("NCHW train-mode", model, x)
run("NCHW train-mode cudnn.benchmark", model, x, True)
mcl = model.to(memory_format=torch.channels_last)
run("channels_last weights, NCHW input", mcl, x, True)
xcl = x.contiguous(memory_format=torch.channels_last)
run("channels_last weights+input", mcl, xcl, True)
mcl.eval()
run("channels_last eval-mode", mcl, xcl, True)
mcl.train()
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    with torch.autocast("cuda", dtype=torch.float16):
        mcl(xcl, return_mine=True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))

# ---- fused forward (msw_gn_act between cuDNN convs)
from minesweeper_ppo_b200.fused_forward import FusedRolloutForward
model.train()
ff = FusedRolloutForward(model)
for _ in range(3):
    ff(x, return_mine=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ff(x, return_mine=True)
b.record(); torch.cuda.synchronize()
print(f"fused forward: {a.elapsed_time(b)/5:.2f} ms / forward")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ff(x, return_mine=True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=90))
