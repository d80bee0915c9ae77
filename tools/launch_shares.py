#!/usr/bin/env python
"""Per-kernel share of the launch list written by `ncu --metrics gpu__time_duration.sum --csv --log-file X`."""
import csv, sys, collections
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = collections.OrderedDict()
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"].split("(")[0]
    t = tot.setdefault(k, [0, 0.0])
    t[0] += 1
    t[1] += float(r["Metric Value"])
s = sum(v[1] for v in tot.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%3d total_ns=%12d share=%5.1f%%" % (k, n, ns, 100 * ns / s))
