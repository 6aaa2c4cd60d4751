"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
idx = {n: i for i, n in enumerate(rows[start])}
agg, tot = collections.OrderedDict(), 0.0
for r in rows[start + 1:]:
    if len(r) < len(idx):
        continue
    v = float(r[idx["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[idx["Metric Unit"]], 1e-3)
    a = agg.setdefault(r[idx["Kernel Name"]][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("%-92s %5s %12s %7s" % ("kernel", "n", "total us", "share"))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%-92s %5d %12.1f %6.1f%%" % (k, n, v, 100 * v / tot))
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
