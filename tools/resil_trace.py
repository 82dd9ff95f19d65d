"""Stage times of one resilience chunk (RHFQ_TRACE=1 python tools/resil_trace.py N [M])."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from reheatfunq_b200.resilience import resilience_run
DP = (2.522017292522833465, 15.37301672440559130, 0.2184765374862465137, 0.2184765374862465137)
MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)
Q = np.linspace(0.99, 0.01, 41)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
M = int(sys.argv[2]) if len(sys.argv) > 2 else 65520
RUNS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
for i in range(RUNS):
    t0 = time.perf_counter()
    res, fail, st = resilience_run(1, MIX, N, M, 100.0, Q, DP, 1.0, 1e-4, 848782, nstreams=64)
    print("run %d: %.3f s  draws %.3f s" % (i, time.perf_counter() - t0, st["host_draw_ns"] * 1e-9), file=sys.stderr, flush=True)
