#!/bin/bash
export TAUV_EXTRA_NVCC="-DTAUV_DEBUG"
python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > /dev/null 2>&1
B=64 DEPTH=1 timeout 200 python tools/mask_trace.py 2>&1 | tail -22
