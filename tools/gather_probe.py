"""Development: minibatch gather at C3 size -- torch index gather of a dense buffer (buffers.py:96-116)
vs msw_gather_encode from bitboard snapshots."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import minesweeper_ppo_b200 as m

N, T = 8192, 128
cfg = m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4)
vec = m.VecMinesweeper(N, cfg, seed=0, api="torch", aux_maps=True)
dense = m.RolloutBuffer(N, T, (10, 16, 16), 256, vec.device, aux_maps=True)
comp = m.CompactRolloutBuffer(vec, T, aux_maps=True)
scratch = vec._alloc_encode()
vec.reset(out=dense.slot(0)); comp.snapshot(0)
for t in range(T):
    nxt = dense.slot(t + 1) if t + 1 < T else scratch
    cur = dense.slot(t)
    vec.step_random(t, out=m.StepOut(obs=nxt.obs, action_mask=nxt.action_mask, rewards=cur.rewards, dones=cur.dones,
                                     mine_labels=nxt.mine_labels, mine_valid=nxt.mine_valid))
    if t + 1 < T:
        comp.snapshot(t + 1)
mb = N * T // 8
order = torch.randperm(N * T, device="cuda")

def timeit(fn, reps=8):
    fn(0); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i % 8)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def dense_gather(i):
    rows = order[i * mb:(i + 1) * mb]
    return dense.obs[rows], dense.action_mask[rows], dense.mine_labels[rows], dense.mine_valid[rows]

def compact_gather(i):
    return comp.gather_obs(order[i * mb:(i + 1) * mb])

td, tc = timeit(dense_gather), timeit(compact_gather)
out_bytes = mb * (46 * 256)
mem = lambda b: sum(t.numel() * t.element_size() for t in b if t is not None) / 1e9
print(json.dumps({
    "minibatch_rows": mb, "dense_index_gather_ms": td, "compact_gather_encode_ms": tc, "speedup": td / tc,
    "compact_write_gbs": out_bytes / tc / 1e6,
    "dense_buffer_gb": mem([dense.obs, dense.action_mask, dense.mine_labels, dense.mine_valid]),
    "compact_buffer_gb": mem([comp.snap_mines, comp.snap_revealed, comp.snap_first])}))
