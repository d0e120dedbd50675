#!/bin/bash
# Final visit of the round: the evidence of tools/profile_round.sh (tests, bench, launch list, decode captures, sweep) plus
# one --set full capture of the round-2 YOLACT / loss kernels.  Everything lands in gpurun_out/.
R=${1:-r2}
bash tools/profile_round.sh $R
cap() {  # cap <kernel regex> <skip> <output name> <command...>
  local k=$1 s=$2 o=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/prof_${o}_$R -f "$@" > gpurun_out/ncu_full_$o.log 2>&1
  echo "ncu $o exit $?"
}
cap mask_umma_kernel 2 mask python tools/mask_once.py
cap mask_umma_kernel 2 maskdepth python tools/mask_depth_once.py
cap scores_tile_kernel 2 scores python tools/detect_probe.py
cap nms_frame_kernel 2 nms python tools/detect_probe.py
EAGER=0 cap ycls_rows_kernel 2 ycls python tools/yolact_loss_once.py
EAGER=0 cap ymask_positive_kernel 2 ymask python tools/yolact_loss_once.py
cap head_pack_kernel 2 heads python tools/heads_once.py
timeout 200 python tools/kp_affinity_once.py 2>&1 | tail -1
timeout 200 python tools/focal_once.py 2>&1 | tail -1
for s in "9 1" "9 2" "15 2"; do timeout 120 python tools/decode_smooth.py $s 2>&1 | tail -1; done
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
