"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 1:]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in data:
    if len(r) <= iv:
        continue
    name = r[ik]
    m = re.search(r"rhfq_kernel<rhfq::(\w+)>", name)
    nm = m.group(1) if m else re.sub(r"\(.*", "", name)[:70]
    v = float(r[iv].replace(",", ""))
    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(r[iu], 1.0)
    agg[nm][0] += 1
    agg[nm][1] += v
    tot += v
print("total %.3f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s n=%4d %10.3f ms %5.1f%%" % (k, n, v, 100 * v / tot))
