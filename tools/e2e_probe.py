"""Where the host-buffer call msw_step_host spends its time at C2 size (65,536 envs): wall-clock per call of the
whole call and of its pieces issued alone (each followed by a stream sync, as the call does)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import minesweeper_ppo_b200 as m

N, K = 65536, 300
dev = torch.device("cuda")
cfg = m.EnvConfig(H=16, W=16, mine_count=40)
vec = m.VecMinesweeper(N, cfg, seed=0, api="torch")
log = torch.empty((K + 8, N), dtype=torch.int32, device=dev)
vec.reset()
for t in range(K + 8):
    vec.step_random(t, actions_out=log[t])
acts = log.cpu().pin_memory()
v = m.VecMinesweeper(N, cfg, seed=0, api="torch")
v.reset()


def wall(fn, n=K):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(5 + i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


us_full = wall(lambda i: v.step_host(acts[i], copy_obs=False, copy_infos=False))
us_full_read = wall(lambda i: int(np.count_nonzero(v.step_host(acts[i], copy_obs=False, copy_infos=False)["done"].numpy())))
d_act = torch.empty((N,), dtype=torch.int32, device=dev)
out = v._alloc_encode()
out.rewards = torch.empty((N,), device=dev); out.dones = torch.empty((N,), dtype=torch.bool, device=dev)


def kernel_only(i):
    v.step(d_act, out=out, want_infos=False)
    torch.cuda.current_stream().synchronize()


d_act.copy_(acts[0])
us_kernel = wall(kernel_only)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100):
    v.step(d_act, out=out, want_infos=False)
e1.record(); torch.cuda.synchronize()
us_kernel_dev = e0.elapsed_time(e1) * 10


def h2d_only(i):
    d_act.copy_(acts[i], non_blocking=True)
    torch.cuda.current_stream().synchronize()


pin_out = torch.empty((N * 5,), dtype=torch.uint8).pin_memory()
dev_out = torch.empty((N * 5,), dtype=torch.uint8, device=dev)


def d2h_only(i):
    pin_out.copy_(dev_out, non_blocking=True)
    torch.cuda.current_stream().synchronize()


def sync_only(i):
    torch.cuda.current_stream().synchronize()


print(f"msw_step_host (actions in, reward+done out, sync)      : {us_full:7.1f} us / call  -> {N / us_full:.1f} M env-steps/s")
print(f"  + the host reads the done flags (np.count_nonzero)    : {us_full_read:7.1f} us / call")
print(f"step launch on device actions + stream sync             : {us_kernel:7.1f} us   (device time of the kernel back to back: {us_kernel_dev:.1f} us)")
print(f"H2D 256 KB + stream sync                                : {wall(h2d_only):7.1f} us")
print(f"D2H 320 KB + stream sync                                : {wall(d2h_only):7.1f} us")
print(f"stream sync on an idle stream                           : {wall(sync_only):7.1f} us")
