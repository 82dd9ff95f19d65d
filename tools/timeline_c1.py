"""Device timeline of one single-posterior construction (config 1) -- where its ~1.5 ms go."""
import sys, os, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
from conftest import gamma_dataset, DEFAULT_PRIOR
from reheatfunq_b200 import PosteriorBatch
from torch.profiler import profile, ProfilerActivity

q, c = gamma_dataset(25, 29189)
for _ in range(5):
    B = PosteriorBatch([[(q, c)]], [[1.0]], DEFAULT_PRIOR); del B
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    B = PosteriorBatch([[(q, c)]], [[1.0]], DEFAULT_PRIOR)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
host = sorted([e for e in prof.events() if e.device_type != torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = min(ev[0].time_range.start, host[0].time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
print("device span %.1f us, busy %.1f us, %d activities; host span %.1f us" % (ev[-1].time_range.end - ev[0].time_range.start, busy, len(ev), host[-1].time_range.end - host[0].time_range.start))
agg = collections.defaultdict(lambda: [0, 0.0])
for h in host:
    agg[h.name][0] += 1; agg[h.name][1] += h.time_range.end - h.time_range.start
print("-- host API time")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-40s n=%4d %9.1f us" % (k[:40], n, t))
print("-- device timeline")
end = ev[0].time_range.start
for e in ev:
    print("%8.1f us dur %6.1f gap %6.1f %s" % (e.time_range.start - t0, e.time_range.end - e.time_range.start, max(0.0, e.time_range.start - end), e.name.split("(")[0].replace("rhfq::", "")[:50]))
    end = max(end, e.time_range.end)
