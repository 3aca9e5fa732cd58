#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
timeout 300 python tools/stream_overlap_probe.py > gpurun_out/r02d_stream_overlap.txt 2>&1; echo "rc=$?" >> gpurun_out/r02d_stream_overlap.txt
timeout 600 python tools/c4_sweep.py > gpurun_out/r02d_c4_sweep_1gpu.jsonl 2> gpurun_out/r02d_c4_sweep.err; echo "rc=$?" >> gpurun_out/r02d_c4_sweep.err
tail -3 gpurun_out/r02d_pytest.log; cat gpurun_out/r02d_stream_overlap.txt; cat gpurun_out/r02d_c4_sweep_1gpu.jsonl | cut -c1-250
