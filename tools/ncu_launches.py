#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): share of GPU time, launches, mean.
Usage: tools/ncu_launches.py launches.csv [series-kernel-substring]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
seq = []
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0][-64:]
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[mu], 1.0)
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += v
    seq.append((name, v))
tot = sum(v for _, v in agg.values())
print(f"total {tot / 1e3:.2f} ms over {len(seq)} launches")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * v / tot:5.1f}%  {v / 1e3:10.2f} ms  n={n:5d}  mean {v / n:9.1f} us  {k}")
if len(sys.argv) > 2:
    ser = [round(v) for n, v in seq if sys.argv[2] in n]
    print(f"{sys.argv[2]} per launch (us):", ser[:60], "..." if len(ser) > 60 else "")
