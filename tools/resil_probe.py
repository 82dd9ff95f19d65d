import sys, time
sys.path.insert(0,'/root/repo')
import numpy as np
from reheatfunq_b200.resilience import resilience_run
DP = (2.522017292522833465, 15.37301672440559130, 0.2184765374862465137, 0.2184765374862465137)
MIX = (31, 3.5, 0.33, 62, 5.5, 0.67)
Q = np.linspace(0.99, 0.01, 41)
for N, M in ((10, 20000), (50, 20000), (100, 20000), (50, 100000)):
    t0 = time.perf_counter()
    res, fail, st = resilience_run(1, MIX, N, M, 100.0, Q, DP, 1.0, 1e-4, 848782, nstreams=8)
    dt = time.perf_counter() - t0
    print("N=%d M=%d: %.2f s -> %.0f sets/s, failures %s, nodes/realisation %.0f, node kernel %.0f ms, draws %.0f ms"
          % (N, M, dt, M/dt, fail, (st['nodes']+st['lz_nodes'])/M, st['node_kernel_ns']*1e-6, st['host_draw_ns']*1e-6))
