"""Measurement for the north-star item "policy forward overlapped with env kernels on CUDA streams".

Three schedules of the same rollout inner loop (forward -> masked sample -> env step) at C3 size
(8,192 envs, medium CNN), timed with CUDA events over STEPS steps:
  serial      one stream, one population of 8,192 envs (what RolloutCollector does);
  env||fwd    two populations of 4,096 envs on two streams, each running its own serial loop, so one
              population's env step / sampler overlaps the other's forward (the north-star's schedule);
  capped      the same two populations with every conv launch capped at half the SMs (max_ctas = 74), so the two
              streams run side by side instead of back to back: an HBM-bound residual layer of one population can
              share the GPU with a tensor-core-bound layer of the other.
(A fourth schedule -- ONE population whose forward is split into two half-batches one layer apart and re-joined
every step -- was measured in round 2 and lost: 4.38 vs 4.04 ms per forward, the one-layer lag is paid every step.)
Prints ms per step of 8,192 envs for each and the env-step / sampler / forward times on their own."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minesweeper_ppo_b200 as m
from minesweeper_ppo_b200.fused_forward import FusedRolloutForward
from minesweeper_ppo_b200.rollout import masked_sample

N, STEPS = 8192, 40
dev = torch.device("cuda")
cfg = m.EnvConfig(H=16, W=16, mine_count=40)
torch.manual_seed(0)
model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                      model_cfg=dict(stem_channels=96, blocks=5, dropout=0.05, value_hidden=256)).to(dev)


class Pop:
    def __init__(self, n, base, max_ctas):
        self.vec = m.VecMinesweeper(n, cfg, seed=0, api="torch", env_id_base=base, aux_maps=True)
        self.fwd = FusedRolloutForward(model, sample_id_base=base, max_ctas=max_ctas)
        self.cur, self.nxt = self.vec._alloc_encode(), self.vec._alloc_encode()
        self.a32 = torch.empty((n,), dtype=torch.int32, device=dev)
        self.r = torch.empty((n,), dtype=torch.float32, device=dev)
        self.d = torch.empty((n,), dtype=torch.bool, device=dev)
        self.vec.reset(out=self.cur)
        self.t = 0

    def step(self):
        logits, values, _ = self.fwd(self.cur.obs, return_mine=True)
        masked_sample(logits, self.cur.action_mask, seed=0, step_index=self.t, actions32=self.a32)
        self.vec.step(self.a32, out=m.StepOut(obs=self.nxt.obs, action_mask=self.nxt.action_mask, rewards=self.r, dones=self.d,
                                              mine_labels=self.nxt.mine_labels, mine_valid=self.nxt.mine_valid), want_infos=False)
        self.cur, self.nxt = self.nxt, self.cur
        self.t += 1


def timed(fn, steps=STEPS):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def two_pops(max_ctas):
    pa, pb = Pop(N // 2, 0, max_ctas), Pop(N // 2, N // 2, max_ctas)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def two():
        with torch.cuda.stream(sa):
            pa.step()
        with torch.cuda.stream(sb):
            pb.step()

    sa.wait_stream(torch.cuda.current_stream()); sb.wait_stream(torch.cuda.current_stream())
    for _ in range(5):
        two()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sa); sb.wait_event(e0)
    for _ in range(STEPS):
        two()
    sa.wait_stream(sb); e1.record(sa)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS


with torch.no_grad():
    p = Pop(N, 0, 0)
    ms_serial = timed(p.step)
    ms_fwd = timed(lambda: p.fwd(p.cur.obs, return_mine=True))
    lg, _, _ = p.fwd(p.cur.obs, return_mine=True)
    ms_sample = timed(lambda: masked_sample(lg, p.cur.action_mask, seed=0, step_index=1, actions32=p.a32))
    ms_env = timed(lambda: p.vec.step(p.a32, out=m.StepOut(obs=p.nxt.obs, action_mask=p.nxt.action_mask, rewards=p.r, dones=p.d,
                                                           mine_labels=p.nxt.mine_labels, mine_valid=p.nxt.mine_valid), want_infos=False))
    ms_serial2 = timed(p.step)
    del p
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    ms_two = two_pops(0)
    ms_cap = {c: two_pops(c) for c in (sms // 2, sms // 2 + 10, sms - 40)}

print(f"components at {N} envs: forward {ms_fwd:.3f} ms, masked sampler {ms_sample * 1e3:.1f} us, env step {ms_env * 1e3:.1f} us")
print(f"serial (one stream)                      : {ms_serial:.3f} ms / step (again after the component timings: {ms_serial2:.3f})")
print(f"env||fwd (two populations, two streams)  : {ms_two:.3f} ms / step  ({100 * (ms_serial / ms_two - 1):+.1f} % steps/s)")
for c, v in ms_cap.items():
    print(f"capped at {c:3d} CTAs per launch            : {v:.3f} ms / step  ({100 * (ms_serial / v - 1):+.1f} % steps/s)")
