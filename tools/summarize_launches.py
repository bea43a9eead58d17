#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and share."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = r[ki].split("(")[0]
    tot[name] += v * scale; cnt[name] += 1
all_us = sum(tot.values())
print(f"| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"| `{k}` | {cnt[k]} | {v/1e3:.2f} | {v/cnt[k]:.1f} | {100*v/all_us:.1f}% |")
print(f"\ntotal {all_us/1e3:.1f} ms over {sum(cnt.values())} launches")
