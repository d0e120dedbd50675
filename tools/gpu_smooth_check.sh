#!/bin/bash
# CenterNet parity tests, then the decode on noise and on smooth maps (box-filtered noise), then the bench's decode line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q > gpurun_out/pytest_cn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cn.log
tail -4 gpurun_out/pytest_cn.log
timeout 120 python tools/decode_once.py 50 2>&1 | tail -1
for s in "5 1" "9 1" "9 2" "15 2"; do timeout 120 python tools/decode_smooth.py $s 2>&1 | tail -1; done
python bench.py --no-yolact --no-cpu-baseline --e2e-steps 4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: value %.0f decode %.1f us (frac %.3f) isolated %.1f us (%.3f) encode %.1f' % (d['value'], d['roofline']['us_per_launch'], d['roofline']['frac'], d['roofline']['us_per_launch_isolated'], d['roofline']['frac_isolated'], d['kernels']['gaussian_encode_us']))"
