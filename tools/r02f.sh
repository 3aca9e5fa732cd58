#!/usr/bin/env bash
# round 2, 2-GPU call: NCCL gradient all-reduce test, bench on 2 ranks (both arms)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_gpus.txt
timeout 600 python -m pytest tests/test_gpu_nccl.py tests/test_shard_gloo.py -m "gpu or not gpu" -q -s > gpurun_out/r02f_nccl_test.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_nccl_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; echo "bench rc=$?" >> gpurun_out/r02f_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/c4_sweep.py 131072 4194304 > gpurun_out/r02f_c4_sweep_2gpu.jsonl 2> gpurun_out/r02f_c4_sweep.err
tail -4 gpurun_out/r02f_nccl_test.txt; tail -2 gpurun_out/r02f_bench_2gpu.err; cat gpurun_out/r02f_c4_sweep_2gpu.jsonl | cut -c1-200
