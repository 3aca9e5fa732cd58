#!/usr/bin/env bash
# round 2, GPU call t: sliced D2H + expansion pipeline, no meta copy; numpy-API tests; e2e legs at the driver's settings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_env.py tests/test_host_expand.py tests/test_late_start.py "tests/test_gpu_reference_live.py::test_lockstep_with_live_reference" tests/test_gpu_reference_live.py::test_unmodified_evaluate_vec_on_cuda_env -q -x > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02t_pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02t_bench20_$i.json 2> gpurun_out/r02t_bench.err; echo "bench rc=$?" >> gpurun_out/r02t_bench.err
done
timeout 600 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02t_bench400.json 2>> gpurun_out/r02t_bench.err
tail -3 gpurun_out/r02t_pytest.log; python - <<'P'
import json
for f in ("r02t_bench20_1", "r02t_bench20_2", "r02t_bench400"):
    d = json.load(open("gpurun_out/%s.json" % f)); print(f, d["value"], d["e2e"]["value"], d["e2e_host_obs"]["value"], d["e2e_host_obs"]["full_rewrite"]["value"]); print(d["e2e"]["step_ms"][:20]); print(d["e2e_host_obs"]["step_ms"][:20])
P
