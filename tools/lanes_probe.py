"""Would two concurrent half-size batches (two host threads, two streams) beat one batch?"""
import sys, os, time, threading
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
from reheatfunq_b200 import PosteriorBatch

R = 10000
q, c, off = bench.make_regions(R, seed=1)
def run(lo, hi, out):
    o = off[lo:hi + 1] - off[lo]
    B = PosteriorBatch.from_packed(q[off[lo]:off[hi]], c[off[lo]:off[hi]], o, np.ones(hi - lo),
                                   np.arange(hi - lo + 1, dtype=np.int64), bench.DEFAULT_PRIOR)
    out.append(B.tail_quantiles(bench.QUANTILES))
for lanes in (1, 2, 3, 4, 1, 2):
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = [[] for _ in range(lanes)]
        bounds = np.linspace(0, R, lanes + 1).astype(int)
        th = [threading.Thread(target=run, args=(bounds[i], bounds[i + 1], outs[i])) for i in range(lanes)]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
    print("lanes %d: %.1f ms" % (lanes, dt))
