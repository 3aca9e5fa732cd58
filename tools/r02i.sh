#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_guard_bands.py "tests/test_gpu_reference_live.py::test_live_collect_rollout_buffer_and_gae" -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
timeout 300 python tools/convgn_probe.py > gpurun_out/r02i_convgn.txt 2>&1; echo "rc=$?" >> gpurun_out/r02i_convgn.txt
timeout 300 python tools/fwd_probe.py > gpurun_out/r02i_fwd.txt 2>&1
tail -4 gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_convgn.txt; head -1 gpurun_out/r02i_fwd.txt
