"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid). Usage: launch_summary.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, gi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:48]
        agg[(name, r[gi])][0] += 1
        agg[(name, r[gi])][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'us total':>12} {'n':>6} {'us each':>9} {'share':>6}  kernel  grid")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1]:12.1f} {v[0]:6d} {v[1] / v[0]:9.2f} {100 * v[1] / tot:5.1f}%  {k[0]}  {k[1]}")
    print(f"{tot:12.1f} total")


if __name__ == "__main__":
    main(sys.argv[1])
