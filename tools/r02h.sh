#!/usr/bin/env bash
# round 2, 8-GPU call: NCCL test at 2 and 8 ranks, bench on 8 ranks, C4 sweep on 8 ranks
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -s > gpurun_out/r02h_nccl_test.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_nccl_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02h_bench_8gpu.json 2> gpurun_out/r02h_bench_8gpu.err; echo "bench rc=$?" >> gpurun_out/r02h_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/c4_sweep.py 524288 1048576 2097152 4194304 > gpurun_out/r02h_c4_sweep_8gpu.jsonl 2> gpurun_out/r02h_c4_sweep.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 tools/c4_sweep.py 262144 4194304 > gpurun_out/r02h_c4_sweep_4gpu.jsonl 2>> gpurun_out/r02h_c4_sweep.err
grep -E "NCCL_OK|passed|failed" gpurun_out/r02h_nccl_test.txt; tail -2 gpurun_out/r02h_bench_8gpu.err; cat gpurun_out/r02h_c4_sweep_8gpu.jsonl gpurun_out/r02h_c4_sweep_4gpu.jsonl | cut -c1-200
