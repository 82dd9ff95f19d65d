"""One warm-up step + one step of the bench workload (for ncu launch lists)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from reheatfunq_b200 import PosteriorBatch

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
q, c, off = bench.make_regions(R, seed=1)
w = np.ones(R)
post_off = np.arange(R + 1, dtype=np.int64)
for i in range(steps):
    t0 = time.perf_counter()
    B = PosteriorBatch.from_packed(q, c, off, w, post_off, bench.DEFAULT_PRIOR)
    t1 = time.perf_counter()
    out = B.tail_quantiles(bench.QUANTILES)
    t2 = time.perf_counter()
    cnt = B.counters()
    print("step %d: create %.1f ms, quantiles %.1f ms, launches %d, node kernel %.2f ms, leaves %d, ok %d"
          % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, cnt["launches"], cnt["node_kernel_ns"] * 1e-6,
             B.saq_leaves(), int((B.status() == 0).sum())))
    del B
