#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, total and share."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    if "gemm_bf16" in name or "paged_attn" in name:
        name += " grid" + r[gi].replace(" ", "")
    try:
        v = float(r[vi])
    except ValueError:
        continue
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v for _, v in agg.values())
print(f"{'kernel':70s} {'n':>6s} {'total_us':>10s} {'avg_us':>8s} {'share':>7s}")
for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:70s} {n:6d} {v / 1e3:10.1f} {v / n / 1e3:8.2f} {100 * v / tot:6.1f}%")
print(f"{'TOTAL':70s} {sum(n for n, _ in agg.values()):6d} {tot / 1e3:10.1f}")
