#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_yolact_gpu.py -m gpu -x -q > gpurun_out/pytest_yl.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_yl.log
tail -6 gpurun_out/pytest_yl.log
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 --steps 20 2>gpurun_out/bench_mask.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print({x: round(k[x],3) for x in ['yolact_scores_us','yolact_detect_us','yolact_detect_frac','mask_us','mask_hbm_frac','mask_depth_us','mask_depth_hbm_frac','mask_binary_nearest_us','mask_binary_bilinear_us','match_anchors_us']})"
tail -2 gpurun_out/bench_mask.err
