#!/usr/bin/env python3
"""Summarise the SASS source page of an ncu report (`ncu -i X.ncu-rep --page source --csv --print-source sass`):
per address range, samples by stall reason, plus the hottest instructions.

usage: ncu_stalls.py src.csv [start_hex end_hex] [--top N]   (addresses relative to the kernel's first instruction)
"""
import csv
import sys


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    rows = list(csv.reader(open(args[0])))
    hdr = rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    data = rows[2:]
    base = min(int(r[col["Address"]], 16) for r in data)
    lo = int(args[1], 16) if len(args) > 1 else 0
    hi = int(args[2], 16) if len(args) > 2 else 1 << 60
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = {k: 0 for k in stall_cols}
    n_samples = 0
    execs = 0
    items = []
    for r in data:
        a = int(r[col["Address"]], 16) - base
        if not (lo <= a < hi):
            continue
        s = int(r[col["# Samples"]] or 0)
        n_samples += s
        execs = max(execs, int(r[col["Instructions Executed"]] or 0))
        for k in stall_cols:
            tot[k] += int(r[col[k]] or 0)
        items.append((s, a, r[col["Source"]].strip()[:70], int(r[col["Instructions Executed"]] or 0)))
    all_samples = sum(int(r[col["# Samples"]] or 0) for r in data)
    print(f"range {lo:#x}-{hi:#x}: samples {n_samples} of {all_samples} ({100.0 * n_samples / max(all_samples, 1):.1f} %), "
          f"max warp-executions of one instruction {execs}")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v:
            print(f"   {k:24s} {v:8d}  {100.0 * v / max(n_samples, 1):5.1f} %")
    top = 12
    if "--top" in sys.argv:
        top = int(sys.argv[sys.argv.index("--top") + 1])
    for s, a, src, ex in sorted(items, reverse=True)[:top]:
        print(f"   {a:05x} {s:7d} samples  x{ex:9d}  {src}")


if __name__ == "__main__":
    main()
