#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q -k "focal" > gpurun_out/pytest_focal.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_focal.log
tail -25 gpurun_out/pytest_focal.log
timeout 200 python tools/focal_once.py 2>&1 | tail -2
