#!/usr/bin/env bash
# round 2, GPU call b: fused conv+GN+residual epilogue (P8 residual stream), host expander
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_env.py tests/test_gpu_guard_bands.py -m gpu -x -q --durations=8 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
timeout 300 python tools/convgn_probe.py > gpurun_out/r02b_convgn.txt 2>&1; echo "rc=$?" >> gpurun_out/r02b_convgn.txt
timeout 300 python tools/fwd_probe.py > gpurun_out/r02b_fwd.txt 2>&1; echo "rc=$?" >> gpurun_out/r02b_fwd.txt
timeout 300 python tools/host_expand_probe.py > gpurun_out/r02b_expand.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c4 --no-train --no-gae > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?" >> gpurun_out/r02b_bench.err
if grep -q "rc=0" gpurun_out/r02b_convgn.txt; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 10 -c 3 -o gpurun_out/r02b_convgn -f python tools/convgn_probe.py 8192 2 > gpurun_out/r02b_ncu.log 2>&1
fi
tail -3 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_convgn.txt; head -3 gpurun_out/r02b_fwd.txt
