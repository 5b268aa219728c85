#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, avg)."""
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        n, t = agg.get(r[ki], (0, 0.0))
        agg[r[ki]] = (n + 1, t + float(r[vi].replace(",", "")) / 1e3)
    tot = sum(t for _, t in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised):", " ".join(sys.argv[2:]))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[:70]:70s} n={n:4d} total={t:9.1f} us share={100 * t / tot:5.1f}% avg={t / n:8.2f} us")


if __name__ == "__main__":
    main()
