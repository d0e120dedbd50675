#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-yolact --no-cpu-baseline --e2e-steps 4 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(json.dumps({k:d[k] for k in ['value','ms_per_step','roofline']},indent=0)[:900])
print(d['kernels'])
PY
