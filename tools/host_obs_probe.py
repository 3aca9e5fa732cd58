"""Per-step time series of msw_step_host with the NumPy result on the host (copy_obs=True) at C2 size, with and
without MADV_HUGEPAGE on the recycled result arrays.  Usage: python tools/host_obs_probe.py [steps]"""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import minesweeper_ppo_b200 as m
from minesweeper_ppo_b200 import env as E

for f in ("enabled", "defrag", "khugepaged/defrag"):
    try:
        print("THP", f, open("/sys/kernel/mm/transparent_hugepage/" + f).read().strip())
    except Exception as e:
        print("THP", f, "?", e)
STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 80
N = 65536
libc = ctypes.CDLL("libc.so.6", use_errno=True)
orig_init = E._ResultSet.__init__


def huge_init(self, n, H, W, sw):
    orig_init(self, n, H, W, sw)
    for a in (self.raw, self.mask, self.shadow):
        lo = (a.ctypes.data + 4095) & ~4095
        ln = (a.ctypes.data + a.nbytes - lo) & ~4095
        rc = libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(ln), 14)      # MADV_HUGEPAGE
        if rc:
            print("madvise failed", ctypes.get_errno())


def run(tag):
    cfg = m.EnvConfig(H=16, W=16, mine_count=40, guarantee_safe_neighborhood=True, step_penalty=1e-4)
    # record a trajectory's actions first (as bench.py does), so nothing else runs between the timed calls
    v = m.VecMinesweeper(N, cfg, seed=0, api="torch")
    v.reset()
    log = torch.empty((STEPS, N), dtype=torch.int32, device=v.device)
    out = v._alloc_encode()
    for t in range(STEPS):
        v.step_random(t, out=out, actions_out=log[t])
    acts = log.cpu().pin_memory()
    del v, out, log
    v = m.VecMinesweeper(N, cfg, seed=0, api="torch")
    v.reset()
    torch.cuda.synchronize()
    ts = []
    pin = None
    for t in range(STEPS):
        t0 = time.perf_counter()
        pin = v.step_host(acts[t], copy_obs=True, copy_infos=False)
        ts.append(1e3 * (time.perf_counter() - t0))
    ts = np.array(ts)
    print(tag, "first two (full writes): %.1f %.1f ms;" % (ts[0], ts[1]),
          "means of steps 2-4 / 5-24 / 25-44 / 45-: %.2f / %.2f / %.2f / %.2f ms" % (ts[2:5].mean(), ts[5:25].mean(), ts[25:45].mean(), ts[45:].mean()))
    print("   series:", " ".join("%.2f" % x for x in ts[2:60]))


run("default arrays      ")
E._ResultSet.__init__ = huge_init
run("MADV_HUGEPAGE arrays")
E._ResultSet.__init__ = orig_init
torch.set_num_threads(1)
run("default, torch 1 thread")
