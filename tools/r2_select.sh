#!/bin/bash
# GPU visit for the block-maxima + select decode: CenterNet parity tests, decode timing, per-kernel launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_centernet_gpu.py -m gpu -x -q > gpurun_out/pytest_cn.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_cn.log
tail -15 gpurun_out/pytest_cn.log
timeout 120 python tools/decode_once.py 50 > gpurun_out/decode_once.log 2>&1; cat gpurun_out/decode_once.log
timeout 120 python tools/decode_once.py 50 64 80 256 256 100 >> gpurun_out/decode_once.log 2>&1; tail -1 gpurun_out/decode_once.log
timeout 120 python tools/decode_once.py 50 64 80 64 64 100 >> gpurun_out/decode_once.log 2>&1; tail -1 gpurun_out/decode_once.log
timeout 120 python tools/decode_once.py 50 1 80 128 128 100 >> gpurun_out/decode_once.log 2>&1; tail -1 gpurun_out/decode_once.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_decode.csv python tools/decode_once.py 5 > gpurun_out/ncu_decode.log 2>&1
grep -E "block_max|select_kernel" gpurun_out/launches_decode.csv | awk -F'","' '{print $5, $(NF)}' | tail -8
