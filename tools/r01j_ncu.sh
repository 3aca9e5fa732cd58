#!/bin/bash
mkdir -p gpurun_out
export MSW_CONV_GN=1
python tools/gn_probe.py > gpurun_out/r01j_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:conv3x3_tc_kernel" -s 9 -c 2 -f -o gpurun_out/r01j_convgn python tools/gn_probe.py > gpurun_out/r01j_ncu.log 2>&1
echo "rc=$?"
