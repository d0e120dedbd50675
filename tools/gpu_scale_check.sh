#!/bin/bash
# scaling check on N GPUs of one box: the bench under torchrun + the mixed workload
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --no-yolact > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_n$N.err; echo "n$N exit $?"; tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_n$N.json').read().strip().splitlines()[-1])
print('N=$N: value %.0f e2e %.0f decode %.1f us frac %.3f h2d/rank %.1f GB/s numa %s cores %s' % (d['value'], d['e2e']['value'], d['kernels']['decode_us'], d['roofline']['frac'], d['e2e']['h2d_gbs_per_rank_min'], d['e2e']['numa_node'], d['e2e']['host_cores_pinned']))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload mixed --steps 20 > gpurun_out/bench_r2_mixed_n$N.json 2> gpurun_out/bench_mixed_n$N.err; echo "mixed n$N exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_mixed_n$N.json').read().strip().splitlines()[-1])
print('mixed N=$N: value %.0f e2e %.0f (%.2f ms/step)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))
PY
nvidia-smi topo -m 2>/dev/null | head -12
