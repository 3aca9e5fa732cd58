#!/usr/bin/env bash
# round 2, GPU call s: per-step series of the host-obs call (THP / warm-up), spin-then-sleep pool, e2e legs
mkdir -p gpurun_out
timeout 300 python tools/host_obs_probe.py 80 > gpurun_out/r02s_host_obs_probe.txt 2>&1
timeout 300 python tests/host_expand_speed.py 65536 0 > gpurun_out/r02s_host_expand_speed.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02s_bench20.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?" >> gpurun_out/r02s_bench.err
cat gpurun_out/r02s_host_obs_probe.txt gpurun_out/r02s_host_expand_speed.txt; python - <<'P'
import json
d = json.load(open("gpurun_out/r02s_bench20.json")); print(d["value"], d["e2e"]["value"], d["e2e_host_obs"]["value"], d["e2e_host_obs"]["full_rewrite"]["value"])
P
