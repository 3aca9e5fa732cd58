"""Development: the env step kernel's launch-shape taper (late CTAs walk 2, then 1 board instead of 3) through the
-DMSW_DEV_KNOBS build: MSW_TAPER_PCT scales the tapered share (0 = every warp walks 3 boards, 100 = the fluid-model
size).  Prints the per-launch kernel time (event pair per launch, mean of 64) at C2 and C4 size."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = os.path.join(ROOT, "tools", "_dev", "libmsw_b200_dev.so")

CHILD = r"""
import os, sys
sys.path.insert(0, %r)
from minesweeper_ppo_b200 import _lib
_lib.LIB_PATH = %r
import numpy as np, torch
import minesweeper_ppo_b200 as m
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
out = []
for name, board, n, ring in ((("C2", (16, 16, 40), 65536, 4),) if os.environ.get("MSW_SWEEP_C2_ONLY") else (("C2", (16, 16, 40), 65536, 4), ("C4", (16, 30, 99), 524288, 2))):
    best = []
    for rep in range(3):
        ms_total, ms_kernel, clocks, _ = bench.time_env_steps(torch, m, dev, 0, 1, torch.cuda.synchronize, board, n, 40, 10, ring)
        best.append(ms_kernel)
    bytes_ = (41 * board[0] * board[1] + 9) * n
    out.append("%%s kernel %%.2f us (%%.0f GB/s; runs %%s)" %% (name, 1e3 * min(best), bytes_ / min(best) / 1e6, " ".join("%%.2f" %% (1e3 * b) for b in best)))
print("bpw=%%s taper_pct=%%s: " %% (os.environ.get("MSW_BPW"), os.environ.get("MSW_TAPER_PCT")) + "   ".join(out), flush=True)
""" % (ROOT, DEV)

if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from minesweeper_ppo_b200 import build as b
    if "--build" in sys.argv or not os.path.exists(DEV):
        os.makedirs(os.path.dirname(DEV), exist_ok=True)
        b.build_dev(DEV)
    if "--build" in sys.argv:
        sys.exit(0)
    combos = [(3, 200), (4, 200), (4, 300), (4, 400), (5, 300), (3, 200)] if "--bpw" in sys.argv else \
             [(3, p_) for p_ in (0, 100, 0, 50, 100, 150, 200, 300, 0)]
    for bpw, pct in combos:
        env = dict(os.environ, MSW_TAPER_PCT=str(pct), MSW_BPW=str(bpw))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output", flush=True)
