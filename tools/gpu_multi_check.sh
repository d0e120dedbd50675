#!/bin/bash
# 2-GPU visit: smoke, the reference arm, the bench under torchrun (weak scaling) and the mixed workload (configs[3])
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?"; tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
print('N=2: value %.0f e2e %.0f decode %.1f us frac %.3f h2d/rank %.1f GB/s' % (d['value'], d['e2e']['value'], d['kernels']['decode_us'], d['roofline']['frac'], d['e2e']['h2d_gbs_per_rank_min']))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload mixed --steps 20 > gpurun_out/bench_r2_mixed_n2.json 2> gpurun_out/bench_mixed_n2.err; echo "mixed n2 exit $?"; tail -2 gpurun_out/bench_mixed_n2.err
cut -c1-700 gpurun_out/bench_r2_mixed_n2.json
timeout 300 python bench.py --workload mixed --steps 20 > gpurun_out/bench_r2_mixed_n1.json 2>/dev/null; cut -c1-300 gpurun_out/bench_r2_mixed_n1.json
