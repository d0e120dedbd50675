#!/bin/bash
# compile-time A/B of the target encode through bench.py, in ONE GPU session
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; TAUV_EXTRA_NVCC="$v" python bench.py --no-yolact --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('value %.0f encode %.1f us (%.3f) decode in-step %.1f us' % (d['value'], k['gaussian_encode_us'], k['gaussian_encode_frac'], k['decode_us']))"
done
