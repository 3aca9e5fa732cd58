"""Development: what bounds msw_conv3x3_gn's no-residual layer (EPI 0)?  Runs the layer at C3 size with parts
switched off through the -DMSW_DEV_KNOBS build (tools/_dev/libmsw_b200_dev.so, built by
minesweeper_ppo_b200.build.build_dev); results are wrong by construction, only the times matter.
  0   everything on                      1  no global stores            2  no statistics barriers
  8   no MMAs (TMA + epilogue only)      16 epilogue = tcgen05.ld + release only (TMA + MMA only)
  32  no horizontal tap shift (every A descriptor atom-aligned)        64 no disable-output-lane masks
Each variant runs in its own process (MSW_CONV_DBG is read at launch time by the dev library)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = os.path.join(ROOT, "tools", "_dev", "libmsw_b200_dev.so")

CHILD = r"""
import os, sys
sys.path.insert(0, %r)
from minesweeper_ppo_b200 import _lib
_lib.LIB_PATH = %r
import torch
from minesweeper_ppo_b200.fused_forward import conv3x3, conv3x3_gn, conv3x3_taps
n, C = 8192, 96
xs = [torch.randn((n, C, 16, 16), device="cuda").half().contiguous(memory_format=torch.channels_last) for _ in range(2)]
rs = [torch.randn((n, 2, 3, 4, 128, 8), device="cuda") for _ in range(2)]
w = (torch.randn((C, C, 3, 3), device="cuda") / (9 * C) ** 0.5).half()
taps = conv3x3_taps(w); bias = 0.1 * torch.randn((C,), device="cuda"); norm = torch.nn.GroupNorm(6, C).cuda()
def timed(fn, reps=20):
    for i in range(4): fn(i & 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i & 1)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps
e0 = timed(lambda i: conv3x3_gn(xs[i], taps, norm, bias, drop_p=0.05, seed=1, call_id=i))
e1 = timed(lambda i: conv3x3_gn(xs[i], taps, norm, bias, res32=rs[i], want32=True))
pl = timed(lambda i: conv3x3(xs[i], taps))
print("dbg=%%s: EPI0 %%.1f us   EPI1 %%.1f us   plain conv %%.1f us" %% (os.environ.get("MSW_CONV_DBG", "0"), e0, e1, pl))
""" % (ROOT, DEV)

if __name__ == "__main__":
    if not os.path.exists(DEV):
        sys.path.insert(0, ROOT)
        from minesweeper_ppo_b200 import build as b
        os.makedirs(os.path.dirname(DEV), exist_ok=True)
        b.build_dev(DEV)
    for pair in (0, 1):
        for dbg in ((0, 1, 16) if "--quick" in sys.argv else (0, 1, 2, 3, 8, 9, 16, 24, 48, 80, 112)):
            env = dict(os.environ, MSW_CONV_DBG=str(dbg), MSW_CONV_PAIR=str(pair))
            try:
                r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
                out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output"
            except subprocess.TimeoutExpired:
                out = "TIMEOUT (hang)"
            print(f"pair={pair} {out}", flush=True)
