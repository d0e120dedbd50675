#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_decode_r1.ncu-rep > profiles/r1_decode_full.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__cluster_max_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]
STALLS = "smsp__average_warps_issue_stalled_"


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    print("(captured with `ncu --set full --clock-control none --import-source on`; times under the profiler are not bench values)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k in WANT:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        rd, wr = (r[hdr.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        u_rd, u_wr = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        print(f"\nDRAM traffic per launch (read + write): {rd} {u_rd} + {wr} {u_wr}\n")
        stalls = sorted(((float(r[i]), h[len(STALLS):].replace("_per_issue_active.ratio", "")) for i, h in enumerate(hdr)
                         if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio") and r[i]), reverse=True)
        print("Warp stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]) + "\n")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[1]
        try:
            isrc, iex, ism = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        except ValueError:
            return
        data = [(r[isrc].strip(), int(r[iex]), int(r[ism])) for r in src[2:] if len(r) > iex and r[iex].isdigit()]
        tot, tots = sum(d[1] for d in data) or 1, sum(d[2] for d in data) or 1
        print(f"Warp instructions executed: {tot}; stall samples: {tots}.  Hottest SASS by stall samples:\n")
        print("| samples % | executed % | SASS |\n|---|---|---|")
        for s, ex, sm in sorted(data, key=lambda d: -d[2])[:12]:
            print(f"| {100 * sm / tots:.1f} | {100 * ex / tot:.2f} | `{s[:80]}` |")
        kinds = {}
        for s, ex, _ in data:
            parts = s.split()
            op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
            op = op.split(".")[0]
            kinds[op] = kinds.get(op, 0) + ex
        print("\nInstruction mix (executed warp instructions): " +
              ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])[:12]))


if __name__ == "__main__":
    main()
