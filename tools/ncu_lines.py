#!/usr/bin/env python3
"""Per-source-line roll-up of an ncu capture (--import-source on, -lineinfo): instructions executed and stall samples
by file:line, grouped into coarse regions.  Usage: python tools/ncu_lines.py REPORT.ncu-rep [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, cur_line, cur_src = None, None, ""
    inst = defaultdict(int); samples = defaultdict(int); tinst = defaultdict(int); text = {}
    header = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            header = r
            i_inst = header.index("Instructions Executed"); i_samp = header.index("# Samples")
            i_tinst = header.index("Thread Instructions Executed")
            continue
        if header is None or len(r) < len(header):
            continue
        if r[0] != "":
            cur_line = (cur_file, int(r[0])); text[cur_line] = r[1].strip()
            continue
        try:
            inst[cur_line] += int(r[i_inst]); samples[cur_line] += int(r[i_samp]); tinst[cur_line] += int(r[i_tinst])
        except ValueError:
            pass
    tot_i = sum(inst.values()) or 1; tot_s = sum(samples.values()) or 1
    print(f"total warp instructions {tot_i}, stall samples {tot_s}")
    print("| file:line | inst % | samples % | avg threads | source |\n|---|---|---|---|---|")
    for k in sorted(inst, key=lambda k: -(inst[k] / tot_i + samples[k] / tot_s))[:top]:
        print(f"| {k[0]}:{k[1]} | {100 * inst[k] / tot_i:.2f} | {100 * samples[k] / tot_s:.2f} | "
              f"{tinst[k] / max(inst[k], 1):.1f} | `{text.get(k, '')[:110]}` |")


if __name__ == "__main__":
    main()
