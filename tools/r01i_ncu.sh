#!/bin/bash
mkdir -p gpurun_out
python tools/gn_probe.py > gpurun_out/r01i_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:conv3x3_tc_kernel|heads_tc_kernel" -s 22 -c 3 -f -o gpurun_out/r01i_tc python tools/gn_probe.py > gpurun_out/r01i_ncu.log 2>&1
echo "tc full rc=$?"
