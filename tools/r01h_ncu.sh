#!/bin/bash
# Round-1 final evidence, profiler passes (each command has run plain first).
mkdir -p gpurun_out
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-rollout --no-gae"
$B > gpurun_out/r01h_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01h_launches.csv $B > gpurun_out/r01h_ncu1.log 2>&1
echo "launch list rc=$?"
E="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-rollout --no-gae --no-e2e"
$E > gpurun_out/r01h_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 10 -c 3 -f -o gpurun_out/r01h_env $E > gpurun_out/r01h_ncu2.log 2>&1
echo "env full rc=$?"
python tools/gn_probe.py > gpurun_out/r01h_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:gn_act_kernel|heads_kernel" -s 24 -c 12 -f -o gpurun_out/r01h_fwd python tools/gn_probe.py > gpurun_out/r01h_ncu3.log 2>&1
echo "fwd full rc=$?"
python tools/gae_probe.py > gpurun_out/r01h_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gae -s 4 -c 2 -f -o gpurun_out/r01h_gae python tools/gae_probe.py > gpurun_out/r01h_ncu4.log 2>&1
echo "gae full rc=$?"
A="python -m pytest tests/test_avoidability.py -m gpu -q -k matches_reference"
$A > gpurun_out/r01h_plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:avoid_kernel -s 4 -c 2 -f -o gpurun_out/r01h_avoid $A > gpurun_out/r01h_ncu5.log 2>&1
echo "avoid full rc=$?"
ls -la gpurun_out/r01h_*
