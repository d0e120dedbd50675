#!/bin/bash
# build variants of the library with different compile-time knobs and time stage 1 with each (run on the GPU box)
for v in "$@"; do
  TAUV_EXTRA_NVCC="$v" python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
  echo "== $v"; python tools/tile_probe.py 2>&1 | head -1; python tools/tile_trace.py 2>&1 | grep -E "rounds per|unit|kernel start"
done
