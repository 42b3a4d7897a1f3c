#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per-kernel totals)."""
import collections
import csv
import sys


def main(path, title, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        tot[row["Kernel Name"]][0] += 1
        tot[row["Kernel Name"]][1] += v
        n += 1
    total = sum(v[1] for v in tot.values())
    print(f"# {title}\n\nTotal {total / 1000:.2f} ms over {n} launches (cold-cache, serialised under ncu: compare SHARES).\n")
    print("| share | total us | launches | us/launch | kernel |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {100 * v[1] / total:.1f}% | {v[1]:.0f} | {v[0]} | {v[1] / v[0]:.1f} | `{k[:100]}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
