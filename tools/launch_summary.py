#!/usr/bin/env python3
"""Aggregate an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) per kernel: launches, total and mean time,
share.  usage: launch_summary.py launches.csv [top]"""
import csv
import collections
import re
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    col = {k: i for i, k in enumerate(hdr)}
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]])
        name = re.sub(r"void |k5::|<unnamed>::|\(anonymous namespace\)::", "", name)
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    print(f"{sum(cnt.values())} launches, {total:.2f} ms under ncu (cold caches, serialised)")
    print(f"{'kernel':70s} {'n':>6s} {'ms':>10s} {'mean us':>10s} {'share':>7s}")
    for name, v in tot.most_common(top):
        print(f"{name[:70]:70s} {cnt[name]:6d} {v:10.3f} {1e3 * v / cnt[name]:10.1f} {100 * v / total:6.1f}%")


if __name__ == "__main__":
    main()
