#!/bin/bash
# phase timeline of select_kernel (debug build): noise, and (SMOOTH="box,passes") smooth maps
mkdir -p gpurun_out
export TAUV_EXTRA_NVCC="-DTAUV_DEBUG"
python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > gpurun_out/build_dbg.log 2>&1 || { tail -5 gpurun_out/build_dbg.log; exit 1; }
timeout 120 python tools/select_trace.py > gpurun_out/select_trace.log 2>&1; cat gpurun_out/select_trace.log
for s in 9,2 15,2; do SMOOTH=$s timeout 120 python tools/select_trace.py 2>&1 | tail -14; done
