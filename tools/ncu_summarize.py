#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full --import-source on) into a markdown summary for profiles/.
Usage: python tools/ncu_summarize.py REPORT.ncu-rep "title / command line" > profiles/NAME.md"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    hdr, units, launches = raw(rep)
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu summary: {title}\n")
    print(f"source report: `{rep}` (ncu --set full --clock-control none --import-source on; cold-cache, serialised replays: "
          "compare shares, not absolutes)\n")
    for row in launches:
        print(f"## {row[col['Kernel Name']]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in col:
                print(f"| {m} | {row[col[m]]} | {units[col[m]]} |")
        print("\nstall reasons (warps stalled per issued instruction, smsp__average_warps_issue_stalled_*_per_issue_active; > 0.2):\n")
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(row[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        for v, h in sorted(stalls, reverse=True):
            if v > 0.2 and h != "selected":
                print(f"* {h}: {v:.2f}")
        extra = ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                 "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
                 "smsp__average_warp_latency_per_inst_issued.ratio"]
        print()
        for m in extra:
            if m in col and row[col[m]] not in ("", "n/a"):
                print(f"* {m}: {row[col[m]]}")
        print()
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) > 2:
        h = rows[1]
        c = {x: i for i, x in enumerate(h)}
        data = [r for r in rows[2:] if len(r) > c["Instructions Executed"] and r[c["Instructions Executed"]]]
        tot = sum(float(r[c["Instructions Executed"]]) for r in data)
        smp = sum(float(r[c["# Samples"]]) for r in data) or 1.0
        print(f"## hottest SASS (by stall samples; total warp instructions {tot:.0f})\n")
        print("| samples % | executed % | avg threads | SASS |\n|---|---|---|---|")
        for r in sorted(data, key=lambda r: -float(r[c["# Samples"]]))[:25]:
            print(f"| {float(r[c['# Samples']]) / smp * 100:.2f} | {float(r[c['Instructions Executed']]) / tot * 100:.2f} | "
                  f"{r[c['Avg. Threads Executed']]} | `{r[1].strip()}` |")


if __name__ == "__main__":
    main()
