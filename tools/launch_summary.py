#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into the table kept under profiles/.

    python tools/launch_summary.py gpurun_out/launches_r1.csv "python bench.py --steps 20 ..." [live_decode_us live_encode_us]
"""
import csv
import sys
from collections import OrderedDict

path, cmd = sys.argv[1], sys.argv[2]
live = [float(x) for x in sys.argv[3:5]]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
head = next(r)
ix = {n: i for i, n in enumerate(head)}
for row in r:
    if len(row) < len(head) or row[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row[ix["Metric Unit"]], 1e-3)
    rows.append((row[ix["Kernel Name"]], row[ix["Grid Size"]], row[ix["Block Size"]], float(row[ix["Metric Value"]].replace(",", "")) * scale))
agg = OrderedDict()
for name, grid, block, us in rows:
    a = agg.setdefault((name, grid, block), [0, 0.0])
    a[0] += 1
    a[1] += us
total = sum(v[1] for v in agg.values())
print(f"# ncu launch list of `{cmd}`\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` — per-launch times are cold-cache and serialised; "
      "the kernels' SHARE of the step is what compares with bench.py's live event timing.\n")
print("| kernel | grid | block | launches | total us | mean us | share of all profiled time |\n|---|---|---|---|---|---|---|")
for (name, grid, block), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:80]}` | {grid} | {block} | {n} | {us:.1f} | {us / n:.2f} | {100 * us / total:.1f}% |")
ours = [(k, v) for k, v in agg.items() if "tauv::" in k[0]]
t_ours = sum(v[1] / v[0] for _, v in ours)
print("\nShare inside the step (our kernels only):\n")
for (name, _, _), (n, us) in sorted(ours, key=lambda kv: -kv[1][1]):
    print(f"- `{name[:60]}`: {100 * (us / n) / t_ours:.1f}% ({us / n:.1f} us per launch under ncu)")
if len(live) == 2:
    s = sum(live)
    print(f"\nLive (bench.py events, same box): decode {live[0]:.1f} us = {100 * live[0] / s:.1f}%, encode {live[1]:.1f} us = {100 * live[1] / s:.1f}%.")
