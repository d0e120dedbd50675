#!/bin/bash
timeout 600 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q 2>&1 | tail -2
python tools/sweep.py 2>/dev/null | head -16
