#!/bin/bash
# Round-1 final evidence, plain runs (no profiler): full GPU test suite, smoke, default bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01h_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r01h_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r01h_smoke.log
python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01h_bench_reference.json 2> gpurun_out/r01h_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r01h_bench.json"))
print("value %.4g  frac %.3f  e2e %.4g  ms/step %.4f" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["ms_per_step"]))
print("gae", {k: d["gae"][k] for k in ("kernel_us", "back_to_back_us", "frac_of_hbm_peak")}, d["gae"]["large"])
print("rollout", d["rollout"]["frames_per_s"], d["rollout"]["ms_rollout"], d["rollout"]["stock_module_forward"]["frames_per_s"])
print("cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
PY
tail -c 600 gpurun_out/r01h_bench_reference.json
