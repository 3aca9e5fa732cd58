#!/usr/bin/env bash
# round 2, 8-GPU call on the final code: NCCL all-reduce test (2 and 8 ranks), bench on 8 ranks (both arms)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02fin8_gpus.txt
timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -s > gpurun_out/r02fin8_nccl_test.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02fin8_nccl_test.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02fin8_bench_reference.json 2> gpurun_out/r02fin8_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02fin8_bench_8gpu.json 2>> gpurun_out/r02fin8_bench.err; echo "bench rc=$?" >> gpurun_out/r02fin8_bench.err
grep -E "NCCL_OK|passed|failed" gpurun_out/r02fin8_nccl_test.txt; tail -2 gpurun_out/r02fin8_bench.err; python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r02fin8_bench_8gpu.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "host_obs", d["e2e_host_obs"]["value"], d["e2e_host_obs"].get("host_threads_per_rank"), "rollout", d["rollout"]["frames_per_s"], "c4", d["c4"]["env_steps_per_s"], "c5", d["train_c5"]["frames_per_s"], d["roofline"]["kernel_ms_per_rank"])
P
