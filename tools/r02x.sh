#!/usr/bin/env bash
# round 2, GPU call x: pipelined GAE kernel: parity tests, A/B against the tile kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gae.py tests/test_gpu_guard_bands.py -q -x > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02x_pytest.log
timeout 600 python tools/gae_ab.py > gpurun_out/r02x_gae_ab.txt 2>&1
tail -3 gpurun_out/r02x_pytest.log; cat gpurun_out/r02x_gae_ab.txt
