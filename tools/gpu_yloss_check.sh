#!/bin/bash
# YOLACT loss: timing against the eager op sequence, then the launch list of one forward + backward
mkdir -p gpurun_out
timeout 500 python tools/yolact_loss_once.py 2>&1 | tail -3
EAGER=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_yloss.csv python tools/yolact_loss_once.py > gpurun_out/ncu_yloss.log 2>&1
python tools/launch_summary.py gpurun_out/launches_yloss.csv "EAGER=0 python tools/yolact_loss_once.py" 2>&1 | grep tauv | head -30
