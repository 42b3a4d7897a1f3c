#!/usr/bin/env python
"""Hot source lines of one kernel in an .ncu-rep (stall samples and executed instructions aggregated per CUDA line from
`ncu --page source --print-source cuda,sass --csv`)."""
import csv
import subprocess
import sys


def main(rep, kernel_regex, top=30):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          f"regex:{kernel_regex}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Warp Stall Sampling (All Samples)" in r)
    hdr = rows[hdr_i]
    si, ss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    # cuda,sass view: every CUDA line row (first column = line number) carries the totals of its SASS rows
    li = hdr.index("Source")
    ss = hdr.index("Warp Stall Sampling (All Samples)")
    agg = {}
    for r in rows[hdr_i + 1:]:
        if len(r) <= ss or not r[0].strip().isdigit():
            continue
        key = (r[0], r[1].strip()[:110])
        v = agg.setdefault(key, [0, 0])
        v[0] += int(r[ss]) if r[ss].isdigit() else 0
        v[1] += int(r[ie]) if r[ie].isdigit() else 0
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    print(f"kernel {kernel_regex}: {tot} stall samples, {toti} warp instructions")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * v[0] / tot:5.1f}% samples {100 * v[1] / toti:5.1f}% inst | {k[0]:>5} {k[1]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
