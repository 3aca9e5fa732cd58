"""Development: msw_gae A/B between the plain-load kernel (gae_kernel, MSW_GAE_TMA=0) and the pipelined TMA kernel
(gae_pipe_kernel, MSW_GAE_TMA=1) through the -DMSW_DEV_KNOBS build.  Per (T, N): cold time (1 GB L2 flush before
every launch, CUDA events around one launch, median of 10), warm time (20 launches replayed from one CUDA graph) and a
SHA-256 of advantages|returns (the kernels must agree bit for bit).  profiles/r02x_gae_ab.txt was made with this
script when the switch (then MSW_GAE_TMA) chose between round 2's first TMA kernel (whole [128 x 32] tiles, commit
9cc9c01) and gae_pipe_kernel."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = os.path.join(ROOT, "tools", "_dev", "libmsw_b200_dev.so")

CHILD = r"""
import hashlib, os, sys
sys.path.insert(0, %r)
from minesweeper_ppo_b200 import _lib
_lib.LIB_PATH = %r
import numpy as np, torch
import minesweeper_ppo_b200 as m
dev = torch.device("cuda", 0)
flush = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for T, N in ((128, 8192), (64, 2048), (128, 65536), (128, 524288), (300, 1024), (33, 4096)):
    g = torch.Generator(device=dev).manual_seed(T * 131 + N)
    buf = m.RolloutBuffer(N, T, (1, 1, 1), 1, dev)
    buf.dones.copy_(torch.rand((T * N,), device=dev, generator=g) < 0.15)
    buf.rewards.copy_(torch.where(buf.dones, torch.tensor(-1.0001, device=dev), torch.tensor(-1e-4, device=dev)))
    buf.values.copy_(0.5 * torch.randn((T * N,), device=dev, generator=g))
    last = 0.5 * torch.randn((N,), device=dev, generator=g)
    cold = []
    for i in range(13):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); buf.compute_gae(last, 0.995, 0.95); b.record(); torch.cuda.synchronize()
        if i >= 3: cold.append(a.elapsed_time(b) * 1e3)
    gr, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        buf.compute_gae(last, 0.995, 0.95); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(20): buf.compute_gae(last, 0.995, 0.95)
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    warm = a.elapsed_time(b) * 1e3 / 20
    h = hashlib.sha256(buf.advantages.cpu().numpy().tobytes() + buf.returns.cpu().numpy().tobytes()).hexdigest()[:16]
    nbytes = 17 * T * N + 4 * N
    print("tma=%%s T=%%d N=%%d: cold %%.2f us  back-to-back %%.2f us (%%.0f GB/s)  sha %%s" %% (os.environ.get("MSW_GAE_TMA"), T, N, float(np.median(cold)), warm, nbytes / warm / 1e3, h), flush=True)
    del buf
""" % (ROOT, DEV)

if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from minesweeper_ppo_b200 import build as b
    if "--build" in sys.argv or not os.path.exists(DEV):
        os.makedirs(os.path.dirname(DEV), exist_ok=True)
        b.build_dev(DEV)
    if "--build" in sys.argv:
        sys.exit(0)
    for pipe in (0, 1):
        env = dict(os.environ, MSW_GAE_TMA=str(pipe))
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
        print(r.stdout.strip(), flush=True)
