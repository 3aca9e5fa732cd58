#!/usr/bin/env bash
# round 2, final code on 2 GPUs (torchrun): bench both arms, NCCL all-reduce test
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02fin2_bench_reference.json 2> gpurun_out/r02fin2_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02fin2_bench_2gpu.json 2>> gpurun_out/r02fin2_bench.err; echo "bench rc=$?" >> gpurun_out/r02fin2_bench.err
tail -2 gpurun_out/r02fin2_bench.err; python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r02fin2_bench_2gpu.json") if l.startswith("{")][-1])
print("value", d["value"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "host_obs", d["e2e_host_obs"]["value"], "rollout", d["rollout"]["frames_per_s"], "c4", d["c4"]["env_steps_per_s"], "c5", d["train_c5"]["frames_per_s"], d["roofline"]["kernel_ms_per_rank"])
P
