#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: share of the total time per kernel.

    python tools/launch_shares.py gpurun_out/launches.csv "<title line>" > profiles/r02_launches_x.txt
"""
import csv
import collections
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1e-3)
    name = re.sub(r"\(.*$", "", r[ik])[:100]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# total %.1f us over %d launches (serialised, cold cache: compare SHARES)" % (total, sum(cnt.values())))
for name, v in sorted(tot.items(), key=lambda kv: -kv[1])[:60]:
    print("%6.2f%% %9.1f us %5d x  %s" % (100 * v / total, v, cnt[name], name))
