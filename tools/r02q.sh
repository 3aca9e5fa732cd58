#!/usr/bin/env bash
# round 2, GPU call q: delta expansion of the NumPy result (msw_host_out.shadow): tests, host throughput on the box, e2e legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env.py tests/test_host_expand.py tests/test_gpu_reference_live.py tests/test_late_start.py -q -x --durations=5 > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log
timeout 300 python tests/host_expand_speed.py 65536 0 > gpurun_out/r02q_host_expand_speed.txt 2>&1
timeout 300 python tests/host_expand_speed.py 65536 1 >> gpurun_out/r02q_host_expand_speed.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02q_bench20.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?" >> gpurun_out/r02q_bench.err
timeout 600 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02q_bench400.json 2>> gpurun_out/r02q_bench.err; echo "bench rc=$?" >> gpurun_out/r02q_bench.err
tail -4 gpurun_out/r02q_pytest.log; cat gpurun_out/r02q_host_expand_speed.txt; python - <<'P'
import json
for f in ("gpurun_out/r02q_bench20.json", "gpurun_out/r02q_bench400.json"):
    try:
        d = json.load(open(f)); print(f, d["value"], d["e2e"]["value"], d["e2e_host_obs"]["value"], d["e2e_host_obs"]["full_rewrite"]["value"])
    except Exception as e: print(f, "ERR", e)
P
