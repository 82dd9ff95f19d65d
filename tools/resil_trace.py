import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from reheatfunq_b200.resilience import resilience_run
DP = (2.522017292522833465, 15.37301672440559130, 0.2184765374862465137, 0.2184765374862465137)
MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)
Q = np.linspace(0.99, 0.01, 41)
for i in range(3):
    t0 = time.perf_counter()
    res, fail, st = resilience_run(1, MIX, 50, 50000, 100.0, Q, DP, 1.0, 1e-4, 848782 + i, nstreams=8)
    print("run %d: %.3f s" % (i, time.perf_counter() - t0), flush=True)
