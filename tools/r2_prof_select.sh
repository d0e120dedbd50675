#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 3 -c 1 -o gpurun_out/prof_select_r2 -f python tools/decode_once.py 5 > gpurun_out/ncu_full_select.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/*.ncu-rep
