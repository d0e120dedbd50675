#!/bin/bash
# compile-time A/B of the mask kernel in one GPU session
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; python tools/mask_trace.py 2>&1 | tail -1; python tools/yolact_probe.py 2>&1 | grep -E "mask_us"
done
