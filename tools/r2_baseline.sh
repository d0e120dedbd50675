#!/bin/bash
# Round-2 GPU visit: parity tests, bench (plain), launch list of the same command.  Output under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench exit $?"
cut -c1-1500 gpurun_out/bench_r2.json; tail -3 gpurun_out/bench_r2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu exit $?"
