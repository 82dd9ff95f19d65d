"""Device timeline of one bench step (CUPTI through torch.profiler): kernel busy time, idle gaps
between kernels and which kernels follow the largest gaps.
    python tools/timeline.py [R] > gpurun_out/timeline.txt"""
import sys, os, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import bench
from reheatfunq_b200 import PosteriorBatch

RESIL = len(sys.argv) > 1 and sys.argv[1] == "resil"
if RESIL:
    # one resilience experiment (config 5, N = 50): python tools/timeline.py resil [M]
    from reheatfunq_b200.resilience import resilience_run
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 65520

    def step():
        return resilience_run(1, bench.MIX, 50, M, 100.0, bench.Q41, bench.DEFAULT_PRIOR, 1.0, 1e-4,
                              848782, nstreams=bench.RESIL_STREAMS)
else:
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    q, c, off = bench.make_regions(R, seed=1)
    w = np.ones(R)
    post_off = np.arange(R + 1, dtype=np.int64)

    def step():
        B = PosteriorBatch.from_packed(q, c, off, w, post_off, bench.DEFAULT_PRIOR)
        out = B.tail_quantiles(bench.QUANTILES)
        del B
        return out


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
busy = 0.0
gaps = collections.defaultdict(lambda: [0, 0.0])
end = ev[0].time_range.start
rows = []
for e in ev:
    s, f = e.time_range.start, e.time_range.end
    gap = max(0.0, s - end)
    name = e.name.split("(")[0].replace("void rhfq::", "").replace("rhfq::", "")[:60]
    gaps[name][0] += 1
    gaps[name][1] += gap
    busy += max(0.0, f - max(s, end))
    end = max(end, f)
    rows.append((s - t0, f - s, gap, name))
print("span %.3f ms, busy %.3f ms, idle %.3f ms, %d device activities" % ((t1 - t0) * 1e-3, busy * 1e-3, (t1 - t0 - busy) * 1e-3, len(ev)))
print("-- idle time in front of (by kernel)")
for k, (n, g) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-62s n=%3d gap %8.1f us" % (k, n, g))
print("-- timeline" + (" (gaps > 20 us only)" if RESIL else ""))
for s, d, g, name in rows:
    if not RESIL or g > 20.0:
        print("%9.1f us  dur %8.1f  gap %6.1f  %s" % (s, d, g, name))

# host side of the large idle gaps: runtime API calls that overlap them
host = [e for e in prof.events() if e.device_type != torch.autograd.DeviceType.CUDA]
host.sort(key=lambda e: e.time_range.start)
print("-- host activity inside device-idle gaps > 40 us")
end = ev[0].time_range.start
for e in ev:
    s, f = e.time_range.start, e.time_range.end
    if s - end > 40.0:
        print("gap %.1f us before %s at %.1f us" % (s - end, e.name[:50], s - t0))
        for h in host:
            if h.time_range.end > end - 20.0 and h.time_range.start < s:
                print("      host %9.1f .. %9.1f us (%7.1f) %s" % (h.time_range.start - t0, h.time_range.end - t0,
                      h.time_range.end - h.time_range.start, h.name[:60]))
    end = max(end, f)
