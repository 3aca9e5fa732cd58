#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rollout.py -k "conv3x3 or fused_forward or dropout_stream" -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
timeout 600 python tools/convgn_ablation.py --quick > gpurun_out/r02k_ablation.txt 2>&1
tail -15 gpurun_out/r02k_pytest.log; cat gpurun_out/r02k_ablation.txt
