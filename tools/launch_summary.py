#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total / mean time, share."""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ni, vi = h.index("Kernel Name"), h.index("Metric Value")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ni])
    name = re.sub(r"^void |madm::|\(anonymous namespace\)::", "", name)[:70]
    t = float(r[vi].replace(",", ""))
    tot[name][0] += 1
    tot[name][1] += t
allt = sum(v[1] for v in tot.values())
print(f"{sum(v[0] for v in tot.values())} launches, {allt / 1e6:.3f} ms (sum of per-launch durations, serialised)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1] / 1e6:9.3f} ms {100 * v[1] / allt:5.1f} %  {v[0]:5d} x {v[1] / v[0] / 1e3:9.1f} us  {k}")
