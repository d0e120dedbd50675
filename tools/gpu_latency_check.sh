#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/latency_probe.py > gpurun_out/latency_r2.log 2>&1; cat gpurun_out/latency_r2.log | tail -12
timeout 200 python tools/latency_yolact.py > gpurun_out/latency_yolact_r2.log 2>&1; tail -6 gpurun_out/latency_yolact_r2.log
timeout 300 python bench.py --workload mixed --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mixed N=1: value %.0f e2e %.0f (%.2f ms/step)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
