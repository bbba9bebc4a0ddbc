"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
tot = 0.0
for r in rows[1:]:
    k = r[ki][:70]
    v = float(r[vi].replace(",", ""))
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print(f"total {tot/1e6:.3f} ms over {len(rows)-1} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{c:5d} {t/1e6:12.3f} ms {100*t/tot:6.1f}%  {k}")
