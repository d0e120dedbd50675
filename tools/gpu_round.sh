#!/bin/bash
# One GPU-box visit: parity tests, probes, bench, launch list.  Output under gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
python tools/tile_probe.py > gpurun_out/tile_probe.log 2>&1; cat gpurun_out/tile_probe.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_tile.csv python tools/tile_once.py > gpurun_out/ncu.log 2>&1
grep -o '"void tauv[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[0-9]*"' gpurun_out/launches_tile.csv | sed 's/"void tauv::\([a-z_0-9]*\).*,"\([0-9]*\)"$/\1 \2 ns/' | tail -12
