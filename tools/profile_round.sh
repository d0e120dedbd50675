#!/bin/bash
# GPU-box visit for the evidence under profiles/: bench (plain), then the ncu launch list of the same command, then one
# --set full capture of each hot kernel.  Everything lands in gpurun_out/.  (Numbers printed under ncu are not bench values.)
R=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err || exit 1
tail -1 gpurun_out/bench_$R.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_cluster_kernel -s 3 -c 1 -o gpurun_out/prof_decode_$R -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gaussian_encode_warp_kernel -s 3 -c 1 -o gpurun_out/prof_encode_$R -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_encode.log 2>&1
python tools/yolact_probe.py > gpurun_out/yolact_probe_$R.json 2> gpurun_out/yolact_probe_$R.err; cat gpurun_out/yolact_probe_$R.json
ncu --set full --clock-control none --import-source on -k regex:mask_umma_kernel -s 2 -c 1 -o gpurun_out/prof_mask_$R -f \
    python tools/mask_once.py > gpurun_out/ncu_full_mask.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scores_tile_kernel -s 2 -c 1 -o gpurun_out/prof_scores_$R -f \
    python tools/yolact_probe.py 16 > gpurun_out/ncu_full_scores.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nms_frame_kernel -s 2 -c 1 -o gpurun_out/prof_nms_$R -f \
    python tools/yolact_probe.py 16 > gpurun_out/ncu_full_nms.log 2>&1
python tools/sweep.py > gpurun_out/sweep_$R.md 2> gpurun_out/sweep.err
ls -la gpurun_out/*.ncu-rep
