#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

  launch list : ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <cmd>
                python tools/ncu_summary.py launches X.csv > profiles/<name>.md
  full capture: ncu --set full -o prof <cmd>;  ncu -i prof.ncu-rep --page raw --csv > raw.csv
                python tools/ncu_summary.py raw raw.csv > profiles/<name>.md
"""
import csv
import sys
from collections import OrderedDict, defaultdict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "smsp__cycles_active.avg", "sm__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def rows_of(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    return list(csv.DictReader(lines))


def launches(path):
    rows = rows_of(path)
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        a = agg.setdefault(name, [0, 0.0, r.get("Grid Size", ""), r.get("Block Size", "")])
        a[0] += 1
        a[1] += us
        total += us
    print(f"# ncu launch list: {path}\n")
    print("per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    print("| kernel | launches | total us | avg us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for name, (n, us, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us:.1f} | {us / n:.1f} | {100 * us / total:.1f}% | {g} | {b} |")
    print(f"\ntotal {total:.1f} us over {sum(a[0] for a in agg.values())} launches")


def raw(path):
    rows = rows_of(path)
    if not rows:
        print("empty")
        return
    # --page raw --csv: one row per kernel launch, one column per metric (first two rows: names, units)
    print(f"# ncu --set full summary: {path}\n")
    hdr = list(rows[0].keys())
    units = rows[0]
    for r in rows[1:]:
        print(f"## {r.get('Kernel Name', '?')}  (ID {r.get('ID')}, grid {r.get('Grid Size')}, block {r.get('Block Size')})\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for k in hdr:
            if any(k == m for m in KEEP) or k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                print(f"| {k} | {r[k]} | {units.get(k, '')} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
