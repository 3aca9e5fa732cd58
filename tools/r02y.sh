#!/usr/bin/env bash
# round 2, GPU call y: pipelined GAE kernel as the only TMA kernel: GAE / buffer / rollout tests, bench gae object,
# ncu --set full captures of both launch shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gae.py tests/test_gpu_guard_bands.py tests/test_gpu_rollout.py tests/test_buffers_reference.py "tests/test_gpu_reference_live.py::test_live_collect_rollout_buffer_and_gae" -q -x > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02y_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-rollout --no-c4 --no-train > gpurun_out/r02y_bench_gae.json 2> gpurun_out/r02y_bench.err; echo "bench rc=$?" >> gpurun_out/r02y_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gae_pipe_kernel -s 3 -c 2 -o gpurun_out/r02y_gae -f python tools/gae_probe.py > gpurun_out/r02y_ncu_gae.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gae_pipe_kernel -s 1 -c 1 -o gpurun_out/r02y_gae_large -f python tools/gae_sizes.py > gpurun_out/r02y_ncu_gae_large.log 2>&1
tail -3 gpurun_out/r02y_pytest.log; python -c "
import json; d=json.load(open('gpurun_out/r02y_bench_gae.json')); print(json.dumps(d['gae'])[:900])"
tail -3 gpurun_out/r02y_ncu_gae.log gpurun_out/r02y_ncu_gae_large.log
