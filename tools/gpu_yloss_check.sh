#!/bin/bash
# YOLACT loss: parity tests, timing against the eager op sequence (138^2 prototypes) and at configs[2]'s 276^2, then the
# launch list of the loss kernels of one forward + backward; smoke()
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python -m pytest tests/test_yolact_gpu.py -m gpu -q -k "loss" 2>&1 | tail -2
EAGER=${EAGER:-0} timeout 500 python tools/yolact_loss_once.py 2>&1 | tail -2
EAGER=0 timeout 500 python tools/yolact_loss_once.py 64 81 32 16 276 550 2>&1 | tail -1
EAGER=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ymask|ycls|ybox|yloss|match_anchors" -c 200 --csv --log-file gpurun_out/launches_yloss.csv python tools/yolact_loss_once.py 64 81 32 16 276 550 > gpurun_out/ncu_yloss.log 2>&1
python tools/launch_summary.py gpurun_out/launches_yloss.csv "EAGER=0 python tools/yolact_loss_once.py 64 81 32 16 276 550 (loss kernels only)" 2>&1 | grep "^| .tauv\|^| .void tauv" | cut -c1-200
