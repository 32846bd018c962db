#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python profiles/tools/launch_summary.py list.csv"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if r is hdr or len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0]
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common():
    print(f"{cnt[k]:5d} launches {v:10.3f} ms total {v / cnt[k]:10.3f} ms/launch {100 * v / s:5.1f}%  {k}")
