#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q > gpurun_out/pytest_cn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cn.log
tail -4 gpurun_out/pytest_cn.log
timeout 120 python tools/decode_once.py 50 2>&1 | tail -1
timeout 120 python tools/decode_once.py 50 64 80 64 64 100 2>&1 | tail -1
python bench.py --no-yolact --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: value %.0f decode %.1f us (frac %.3f) isolated %.1f us (%.3f) encode %.1f' % (d['value'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['roofline']['us_per_launch_isolated'], d['roofline']['frac_isolated'], d['kernels']['gaussian_encode_us']))"
export TAUV_EXTRA_NVCC="-DTAUV_DEBUG"
python -c "import tauv_vision_b200 as tv; tv.build(force=True)" > gpurun_out/build_dbg.log 2>&1 || { tail -5 gpurun_out/build_dbg.log; exit 1; }
timeout 120 python tools/select_trace.py > gpurun_out/select_trace.log 2>&1; cat gpurun_out/select_trace.log
