#!/bin/bash
# compile-time A/B of the fused mask + depth kernel (and the mask writer) in one GPU session
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; B=64 DEPTH=1 python tools/mask_trace.py 2>&1 | grep -E "means"; python tools/mask_depth_once.py 2>&1 | tail -1
done
