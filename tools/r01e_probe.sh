#!/bin/bash
# One GPU call: parity tests, GAE / forward timings, e2e transfer-plan comparison.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r01e_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r01e_pytest.log
python tools/gae_probe.py > gpurun_out/r01e_gae.json 2> gpurun_out/r01e_gae.err; cat gpurun_out/r01e_gae.json
python tools/fwd_probe.py > gpurun_out/r01e_fwd.txt 2>&1; head -3 gpurun_out/r01e_fwd.txt
for mode in 0 1 2 3; do
  MSW_HOST_MODE=$mode python bench.py --steps 400 --warmup 20 --no-gae --no-rollout --no-cpu-baseline \
      > gpurun_out/r01e_e2e_mode$mode.json 2> gpurun_out/r01e_e2e_mode$mode.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r01e_e2e_mode$mode.json"))
print("mode $mode value %.4g e2e %.4g us/step %.1f host_obs %.4g" % (d["value"], d["e2e"]["value"], 65536 / d["e2e"]["value"] * 1e6, d["e2e_host_obs"]["value"]))
PY
done
