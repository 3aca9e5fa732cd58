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
# the C-ABI call alone (prepared argument structs, no Python mirror work around it)
_key = (False, False, v.aux_maps)
_io, _h, _rd, _rs, _rio, _rh = v._host_calls[_key]
_stream = torch.cuda.current_stream().cuda_stream
_ptrs = [acts[i].data_ptr() for i in range(K + 8)]
us_c = wall(lambda i: v._L.msw_step_host(_rd, _rs, _rio, _ptrs[i], _rh, N, _stream))
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


# the same three operations captured once in a CUDA graph (fixed pinned source / destination) and replayed: what a
# graph-replaying msw_step_host could reach for a caller that reuses ONE pinned action buffer (built, measured at
# 135 us against 141 us, no gain for rotating buffers, removed: profiles/experiments/r02w_step_host_graph_replay.patch)
stage_in = torch.empty((N,), dtype=torch.int32).pin_memory()
gout = v._alloc_encode()
gout.rewards = dev_out[:N * 4].view(torch.float32); gout.dones = dev_out[N * 4:].view(torch.bool)
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        d_act.copy_(stage_in, non_blocking=True); v.step(d_act, out=gout, want_infos=False); pin_out.copy_(dev_out, non_blocking=True)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        d_act.copy_(stage_in, non_blocking=True); v.step(d_act, out=gout, want_infos=False); pin_out.copy_(dev_out, non_blocking=True)


def graph_step(i):
    g.replay()
    side.synchronize()


stage_in.copy_(acts[0])
us_graph = wall(graph_step)
print(f"msw_step_host (actions in, reward+done out, sync)      : {us_full:7.1f} us / call  -> {N / us_full:.1f} M env-steps/s")
print(f"  the C-ABI call alone (ctypes, prepared structs)        : {us_c:7.1f} us / call")
print(f"  H2D + step + D2H captured in one CUDA graph, replay + sync: {us_graph:7.1f} us / call (fixed action: mostly no-op clicks)")
print(f"  + the host reads the done flags (np.count_nonzero)    : {us_full_read:7.1f} us / call")
print(f"step launch on device actions + stream sync             : {us_kernel:7.1f} us   (device time of the kernel back to back: {us_kernel_dev:.1f} us)")
print(f"H2D 256 KB + stream sync                                : {wall(h2d_only):7.1f} us")
print(f"D2H 320 KB + stream sync                                : {wall(d2h_only):7.1f} us")
print(f"stream sync on an idle stream                           : {wall(sync_only):7.1f} us")
