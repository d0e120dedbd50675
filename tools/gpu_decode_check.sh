#!/bin/bash
timeout 600 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/sweep.py 2>/dev/null | head -16
python bench.py --no-yolact --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench in-step %.1f us (%.3f) isolated %.1f us value %.0f' % (d['kernels']['decode_us'], d['roofline']['frac'], d['kernels']['decode_isolated_us'], d['value']))"
