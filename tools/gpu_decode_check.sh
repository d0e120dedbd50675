#!/bin/bash
# CenterNet parity tests, the CenterNet half of the sweep, the focal-loss timing and the bench's decode / encode numbers
timeout 600 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/sweep.py 2>/dev/null | head -16
timeout 200 python tools/focal_once.py 2>&1 | tail -1
python bench.py --no-yolact --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('bench: value %.0f decode in-step %.1f us (%.3f) isolated %.1f us encode %.1f us (%.3f)' % (d['value'], k['decode_us'], d['roofline']['frac'], k['decode_isolated_us'], k['gaussian_encode_us'], k['gaussian_encode_frac']))"
