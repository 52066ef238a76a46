#!/usr/bin/env python3
"""Roll an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file LIST.csv <command>)
up by kernel: launches, total and mean duration, share.  Usage: python tools/launch_list.py LIST.csv  ->  markdown table."""
import collections
import csv
import re
import sys


def short_name(name: str) -> str:
    m = re.search(r"(k_\w+)(<[^>]*>)?", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    if "FillFunctor" in name:
        return "torch fill (" + ("L2 flush, uint8" if "unsigned char" in name else "zero / constant") + ")"
    if "ncclDevKernel" in name:
        return re.search(r"ncclDevKernel\w+", name).group(0)
    return re.sub(r"\(.*", "", name)[:70]


def main():
    with open(sys.argv[1]) as f:
        lines = [line for line in f if line.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        a = agg.setdefault(short_name(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"])
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e6:.3f} | {100 * v[1] / total:.1f} % |")


if __name__ == "__main__":
    main()
