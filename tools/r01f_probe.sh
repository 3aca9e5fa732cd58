#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_guard_bands.py -m gpu -x -q > gpurun_out/r01f_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r01f_pytest.log
python tools/fwd_probe.py > gpurun_out/r01f_fwd.txt 2>&1; head -1 gpurun_out/r01f_fwd.txt; cut -c1-100,190-260 gpurun_out/r01f_fwd.txt | sed -n 5,16p
