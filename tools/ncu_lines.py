"""Instruction / stall-sample share per CUDA source line of one result of an .ncu-rep
(ncu -i rep --page source --print-source cuda,sass --kernel-id :::N), plus the SASS opcode mix."""
import collections
import csv
import re
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
ops, osmp = collections.Counter(), collections.Counter()
tot = ts = 0
name = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        name = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        iex = hdr.index("Instructions Executed")
        ismp = hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr is None or len(r) <= max(iex, ismp):
        continue
    try:
        n, s = int(r[iex]), int(r[ismp])
    except ValueError:
        continue
    if r[0] != "":
        k = (cur, int(r[0]))
        agg[k][0] += n
        agg[k][1] += s
        agg[k][2] = r[1].strip()[:100]
        tot += n
        ts += s
    else:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3].strip())
        op = m.group(2).split(".")[0] if m else r[3][:10]
        ops[op] += n
        osmp[op] += s
print("kernel %s: %d warp instructions, %d stall samples" % (name[:80], tot, ts))
print("-- opcode mix")
t2 = sum(ops.values()) or 1
s2 = sum(osmp.values()) or 1
for op, n in ops.most_common(16):
    print("%-10s %6.2f%% instr %6.2f%% samples" % (op, 100.0 * n / t2, 100.0 * osmp[op] / s2))
print("-- source lines")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-20s %5d %5.2f%% instr %5.2f%% smp  %s" % (k[0][:20], k[1], 100.0 * v[0] / max(tot, 1),
                                                      100.0 * v[1] / max(ts, 1), v[2]))
