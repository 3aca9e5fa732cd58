#!/usr/bin/env bash
# round 2, GPU call zz (after the e2e measurement fixes): the whole evidence set on the final code: full GPU suite, smoke, bench (both arms), ncu launch
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r02zz_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02zz_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zz_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02zz_smoke.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02zz_bench_reference.json 2> gpurun_out/r02zz_bench.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02zz_bench.json 2>> gpurun_out/r02zz_bench.err; echo "bench rc=$?" >> gpurun_out/r02zz_bench.err
if grep -q "bench rc=0" gpurun_out/r02zz_bench.err; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02zz_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/r02zz_ncu_bench.log 2>&1
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gae_pipe_kernel -s 86 -c 1 -o gpurun_out/r02zz_gae_large -f python tools/gae_sizes.py > gpurun_out/r02zz_ncu_gae_large.log 2>&1
