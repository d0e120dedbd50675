#!/bin/bash
# build variants of the library with different compile-time knobs and time stage 1 with each (run on the GPU box)
for v in "$@"; do
  IFS=, read -r w s <<< "$v"
  TAUV_EXTRA_NVCC="-DTAUV_ROUND_W=$w -DTAUV_SVC_ITEMS=$s" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== ROUND_W=$w SVC_ITEMS=$s"; python tools/tile_probe.py 2>&1 | head -1; python tools/tile_trace.py 2>&1 | grep -E "service:|rounds per|unit"
done
