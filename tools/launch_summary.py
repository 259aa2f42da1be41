#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised launches:
compare SHARES, not absolutes).   usage: launch_summary.py launches.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void rambl::<unnamed>::", "").replace("rambl::<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(r[ui], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
print("%-44s %8s %12s %7s" % ("kernel", "launches", "device ms", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %8d %12.3f %6.1f%%" % (k[:44], a[0], a[1], 100 * a[1] / tot))
print("%-44s %8d %12.3f" % ("total", sum(a[0] for a in agg.values()), tot))
