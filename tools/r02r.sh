#!/usr/bin/env bash
# round 2, GPU call r: full GPU suite + full bench line at the driver's settings (pre-roll, delta expansion, lazy infos)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?" >> gpurun_out/r02r_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02r_bench_reference.json 2>> gpurun_out/r02r_bench.err
timeout 300 python - > gpurun_out/r02r_numpy_step.txt 2>&1 <<'P'
import time, numpy as np, torch
import minesweeper_ppo_b200 as m
N = 65536
v = m.VecMinesweeper(N, m.EnvConfig(H=16, W=16, mine_count=40, step_penalty=1e-4), seed=0)
b = v.reset()
rng = np.random.default_rng(0)
ts = []
for t in range(40):
    s = rng.random(b["action_mask"].shape, dtype=np.float32); s[~b["action_mask"]] = -1; a = s.argmax(1)
    t0 = time.perf_counter(); b, r, d, infos = v.step(a); ts.append(time.perf_counter() - t0)
print("VecMinesweeper.step (api=numpy, the reference call), 65,536 envs: median %.2f ms/step = %.3g env-steps/s; infos['aux'][7] = %r"
      % (1e3 * np.median(ts[10:]), N / np.median(ts[10:]), infos["aux"][7]))
t0 = time.perf_counter(); lst = list(infos["aux"]); print("materialising all 65,536 aux dicts: %.1f ms" % (1e3 * (time.perf_counter() - t0)))
P
tail -3 gpurun_out/r02r_pytest.log; cat gpurun_out/r02r_numpy_step.txt; python - <<'P'
import json
d = json.load(open("gpurun_out/r02r_bench.json")); print(d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_host_obs"]["value"], d["e2e_host_obs"]["full_rewrite"]["value"], d["rollout"]["frames_per_s"], d["c4"]["env_steps_per_s"], d["train_c5"]["frames_per_s"])
P
