#!/bin/bash
# GPU-box visit for the evidence under profiles/: bench (plain), then the ncu launch list of the same command, then one
# --set full capture of the decode kernel and of the encode kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err || exit 1
tail -1 gpurun_out/bench_r1.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_cluster_kernel -s 3 -c 1 -o gpurun_out/prof_decode_r1 -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_decode.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gaussian_encode_warp_kernel -s 3 -c 1 -o gpurun_out/prof_encode_r1 -f \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_encode.log 2>&1
python tools/yolact_probe.py > gpurun_out/yolact_probe_r1.json 2> gpurun_out/yolact_probe_r1.err; cat gpurun_out/yolact_probe_r1.json
ls -la gpurun_out/*.ncu-rep
