#!/usr/bin/env bash
# round 2, GPU call e: full GPU suite, EPI0 ncu capture, bench launch list, full bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02e_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02e_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 4 -c 1 -o gpurun_out/r02e_convgn_epi0 -f python tools/convgn_probe.py 8192 2 > gpurun_out/r02e_ncu.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?" >> gpurun_out/r02e_bench.err
if grep -q "rc=0" gpurun_out/r02e_bench.err; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02e_ncu_bench.log 2>&1
fi
tail -5 gpurun_out/r02e_pytest.log; cat gpurun_out/r02e_smoke.log
