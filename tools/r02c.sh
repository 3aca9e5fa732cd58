#!/usr/bin/env bash
# round 2, GPU call c: two-stream forward, guard / property tests, e2e fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_floodfill_property.py tests/test_gpu_env.py -m gpu -x -q --durations=5 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
OVERLAP=0 timeout 300 python tools/fwd_probe.py 2>&1 | head -3 > gpurun_out/r02c_fwd_serial.txt
OVERLAP=1 timeout 300 python tools/fwd_probe.py > gpurun_out/r02c_fwd_overlap.txt 2>&1
timeout 300 python tools/stream_overlap_probe.py > gpurun_out/r02c_stream_overlap.txt 2>&1; echo "rc=$?" >> gpurun_out/r02c_stream_overlap.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c4 --no-train --no-gae > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?" >> gpurun_out/r02c_bench.err
tail -4 gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_fwd_serial.txt; head -3 gpurun_out/r02c_fwd_overlap.txt; cat gpurun_out/r02c_stream_overlap.txt
