#!/usr/bin/env bash
# round 2, first GPU call: live-reference tests, full GPU suite, bench both arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt 2>&1
lscpu | head -20 >> gpurun_out/r02a_smi.txt
ls baseline/_ref > gpurun_out/r02a_ref_ls.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?" >> gpurun_out/r02a_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_reference.json 2> gpurun_out/r02a_bench_reference.err
tail -5 gpurun_out/r02a_pytest.log
