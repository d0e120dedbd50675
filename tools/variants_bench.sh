#!/bin/bash
# build variants of the library (compile-time knobs) and run the bench with each, in ONE GPU session (boxes differ)
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; python bench.py --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['kernels']['decode_us'])"
done
