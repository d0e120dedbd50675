#!/bin/bash
# compile-time A/B of the mask kernels through bench.py's configs[2] timings, in ONE GPU session
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; TAUV_EXTRA_NVCC="$v" timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print({x: round(k[x],3) for x in ['mask_us','mask_hbm_frac','mask_depth_us','mask_depth_hbm_frac']})"
done
