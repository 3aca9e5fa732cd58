#!/usr/bin/env bash
# round 2, final evidence set (after the launch-shape taper): full GPU suite, smoke, bench (both arms), ncu --set full of
# the env step kernel from the bench command, ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02fin_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02fin_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02fin_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02fin_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02fin_bench_reference.json 2> gpurun_out/r02fin_bench.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02fin_bench.json 2>> gpurun_out/r02fin_bench.err; echo "bench rc=$?" >> gpurun_out/r02fin_bench.err
if grep -q "bench rc=0" gpurun_out/r02fin_bench.err; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 40 -c 3 -o gpurun_out/r02fin_env -f python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-gae --no-rollout --no-c4 --no-train > gpurun_out/r02fin_ncu_env.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02fin_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02fin_ncu_bench.log 2>&1
fi
tail -3 gpurun_out/r02fin_pytest.log; tail -2 gpurun_out/r02fin_smoke.log; cat gpurun_out/r02fin_bench.err
