"""DRAM traffic per kernel from an ncu pass with --metrics dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum (csv log) -> profiles/rNN_traffic.json, keyed by the
kernel names bench.py reports in its roofline objects."""
import collections
import csv
import json
import re
import sys

src, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]
ik, im, iv, iu = (h.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
agg = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    m = re.search(r"rhfq_kernel<rhfq::(\w+)>", r[ik])
    nm = m.group(1) if m else re.sub(r"\(.*", "", r[ik]).replace("void ", "")
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if "byte" in u.lower():
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
    else:
        v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1)
    agg[nm][r[im]] += v
    if r[im] == "gpu__time_duration.sum":
        cnt[nm] += 1


def entry(names):
    b = sum(agg[n]["dram__bytes_read.sum"] + agg[n]["dram__bytes_write.sum"] for n in names)
    t = sum(agg[n]["gpu__time_duration.sum"] for n in names)
    n = sum(cnt[n] for n in names)
    return {"bytes_per_launch": b / max(n, 1), "launches": n,
            "dram_GBps_under_ncu": (b / 1e9 / (t * 1e-3)) if t > 0 else None,
            "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, %s" % src}


res = {"rhfq_saq_post": entry([n for n in agg if n.startswith("rhfq_saq_post") or n.startswith("rhfq::rhfq_saq_post")]),
       "rhfq_saq_step_bli": entry([n for n in agg if "rhfq_saq_step_bli" in n]),
       "rhfq_node_quad + rhfq_norm_quad": entry([n for n in agg if n.split("::")[-1] in ("rhfq_node_quad", "rhfq_norm_quad")])}
res = {k: v for k, v in res.items() if v["launches"] > 0}
json.dump(res, open(out, "w"), indent=1)
for k, v in res.items():
    print(k, v)
for nm, d in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])[:14]:
    t = d["gpu__time_duration.sum"]
    b = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    print("%-44s n=%3d %8.3f ms %8.3f GB %7.1f GB/s" % (nm[:44], cnt[nm], t, b / 1e9, b / 1e9 / max(t * 1e-3, 1e-12)))
