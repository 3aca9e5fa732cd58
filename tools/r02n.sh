#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
timeout 300 python tools/fwd_probe.py > gpurun_out/r02n_fwd.txt 2>&1
timeout 300 python tools/convgn_probe.py > gpurun_out/r02n_convgn.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c4 --no-train > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?" >> gpurun_out/r02n_bench.err
tail -4 gpurun_out/r02n_pytest.log; head -1 gpurun_out/r02n_fwd.txt; cat gpurun_out/r02n_convgn.txt
