#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/match_once.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:match_anchors_kernel -s 2 -c 1 -o gpurun_out/prof_match_r2 -f python tools/match_once.py > gpurun_out/ncu_match.log 2>&1; echo "ncu exit $?"
