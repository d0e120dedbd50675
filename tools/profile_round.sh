#!/bin/bash
# GPU-box visit for the evidence under profiles/: parity tests, bench (plain), then the ncu launch list of the same
# command, then one --set full capture of each hot kernel.  Everything lands in gpurun_out/.  (Numbers printed under ncu
# are not bench values.)
R=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$R.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$R.log; tail -2 gpurun_out/pytest_$R.log
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err || { tail -5 gpurun_out/bench_$R.err; exit 1; }
tail -1 gpurun_out/bench_$R.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:block_max_kernel -s 3 -c 1 -o gpurun_out/prof_blockmax_$R -f \
    python tools/decode_once.py 5 > gpurun_out/ncu_full_blockmax.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 3 -c 1 -o gpurun_out/prof_select_$R -f \
    python tools/decode_once.py 5 > gpurun_out/ncu_full_select.log 2>&1
python tools/sweep.py > gpurun_out/sweep_$R.md 2> gpurun_out/sweep.err; tail -3 gpurun_out/sweep.err
ls -la gpurun_out/*.ncu-rep
