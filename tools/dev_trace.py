"""Repeated device-resident steps with RHFQ_TRACE=1 (find the stage behind step-time spikes)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench
from reheatfunq_b200 import PosteriorBatch

R = 10000
q, c, off = bench.make_regions(R, seed=1)
dev = torch.device("cuda:0")
dq = torch.tensor(q, device=dev); dc = torch.tensor(c, device=dev)
doff = torch.tensor(off, device=dev); dw = torch.ones(R, dtype=torch.float64, device=dev)
dpost = torch.arange(R + 1, dtype=torch.int64, device=dev)
dprior = torch.tensor(bench.DEFAULT_PRIOR, dtype=torch.float64, device=dev)
nq = len(bench.QUANTILES)
dt = torch.tensor(np.tile(bench.QUANTILES, R), device=dev)
dtoff = torch.arange(0, R * nq + 1, nq, dtype=torch.int64, device=dev)
dout = torch.empty(R * nq, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    flush.fill_(i & 255)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    B = PosteriorBatch.from_device(dq.data_ptr(), dc.data_ptr(), doff.data_ptr(), dw.data_ptr(),
                                   dpost.data_ptr(), R, dprior.data_ptr(), dw.numel(), dq.numel(),
                                   device=0)
    B.query_device("tail_quantiles", dt.data_ptr(), dtoff.data_ptr(), dout.data_ptr(), dout.numel())
    torch.cuda.synchronize()
    print("STEP %d %.1f ms" % (i, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
    del B
