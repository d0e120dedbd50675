#!/bin/bash
# build variants of the library (compile-time knobs) and run the bench with each, in ONE GPU session (boxes differ)
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; TAUV_EXTRA_NVCC="$v" python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-yolact --e2e-steps 2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f in-step %.1f us (%.3f) isolated %.1f us' % (d['value'], d['kernels']['decode_us'], d['roofline']['frac'], d['kernels']['decode_isolated_us']))"
done
