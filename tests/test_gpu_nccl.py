"""-m gpu, needs >= 2 GPUs: the ONE collective of the system -- the flat gradient all-reduce of the PPO update
(BASELINE configs[4]; SURVEY 4.7 "NCCL allreduce smoke 2/4/8 ranks") -- over NCCL, one process per GPU.
Replicas must stay bit-identical over three optimizer steps on different per-rank data.  On a one-GPU
box the test reports itself skipped (two NCCL ranks cannot share a device); run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu` (log: profiles/r02_nccl_test.txt)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, hashlib
sys.path.insert(0, os.environ["MSW_ROOT"])
import torch, torch.distributed as dist
import minesweeper_ppo_b200 as m
from minesweeper_ppo_b200.train import FlatGradAllReduce, PPOConfig, ppo_update
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = m.build_model("cnn_residual", obs_shape=(10, 16, 16),
                      model_cfg=dict(stem_channels=96, blocks=2, dropout=0.0, value_hidden=64)).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
grads = FlatGradAllReduce(model, time_it=True)
cfg = m.EnvConfig(H=16, W=16, mine_count=40)
vec = m.VecMinesweeper(256, cfg, seed=0, api="torch", env_id_base=256 * rank)
col = m.RolloutCollector(vec, 8, aux_maps=True, sample_seed=0)
pc = PPOConfig(aux_mine_weight=0.2)
for step in range(3):
    buf, aux = col.collect(model)
    buf.compute_gae(aux["last_values"], 0.995, 0.95)
    for batch in buf.get_minibatches(1024):
        ppo_update(model, opt, batch, pc, None, grads, want_stats=False)
        break
torch.cuda.synchronize()
h = hashlib.sha256()
for p in model.parameters():
    h.update(p.detach().cpu().numpy().tobytes())
digest = h.hexdigest()
out = [None] * world
dist.all_gather_object(out, digest)
ms = grads.allreduce_ms()
if rank == 0:
    assert len(set(out)) == 1, out
    assert ms is not None and ms["calls"] == 3
    print("NCCL_OK ranks=%d bucket_bytes=%d allreduce_mean_ms=%.3f digest=%s" % (world, grads.numel * 4, ms["mean_ms"], digest[:16]))
dist.destroy_process_group()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 8])
def test_flat_grad_allreduce_nccl(world, tmp_path):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (NCCL ranks cannot share a device); this box has {torch.cuda.device_count()}")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MSW_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and "NCCL_OK" in r.stdout
