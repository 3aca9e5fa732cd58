#!/usr/bin/env bash
# round 2, GPU call v: msw_step_host as one replayed CUDA graph: tests that go through it, probe, e2e legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env.py tests/test_gpu_rollout.py tests/test_late_start.py tests/test_gpu_reference_live.py -q -x > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02v_pytest.log
timeout 300 python tools/e2e_probe.py > gpurun_out/r02v_e2e_probe.txt 2>&1
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02v_bench20_$i.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?" >> gpurun_out/r02v_bench.err
done
timeout 600 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02v_bench400.json 2>> gpurun_out/r02v_bench.err
tail -3 gpurun_out/r02v_pytest.log; cat gpurun_out/r02v_e2e_probe.txt; python - <<'P'
import json
for f in ("r02v_bench20_1", "r02v_bench20_2", "r02v_bench400"):
    d = json.load(open("gpurun_out/%s.json" % f)); print(f, d["value"], d["e2e"]["value"], d["e2e_host_obs"]["value"]); print(d["e2e"]["step_ms"][:20])
P
