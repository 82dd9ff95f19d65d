"""Single-posterior latency (BASELINE.json config 1: N = 25, pdf + cdf + tail on 1000 P_H nodes):
what HeatFlowAnomalyPosterior.__init__ + three evaluations pay (posterior.py:361-368)."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from conftest import gamma_dataset, DEFAULT_PRIOR
from reheatfunq_b200 import PosteriorBatch

q, c = gamma_dataset(25, 29189)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
t_create, t_eval, launches = [], [], 0
for i in range(reps + 5):
    t0 = time.perf_counter()
    B = PosteriorBatch([[(q, c)]], [[1.0]], DEFAULT_PRIOR)
    t1 = time.perf_counter()
    x = np.linspace(0, B.Qmax()[0], 1000)
    B.pdf(x); B.cdf(x); B.tail(x)
    t2 = time.perf_counter()
    if i >= 5:
        t_create.append((t1 - t0) * 1e3); t_eval.append((t2 - t1) * 1e3)
    launches = B.counters()["launches"]
    del B
print(json.dumps(dict(config="C1 single posterior N=25", create_ms_median=float(np.median(t_create)),
                      eval_pdf_cdf_tail_1000_ms_median=float(np.median(t_eval)),
                      total_ms_median=float(np.median(np.add(t_create, t_eval))), launches=int(launches))))
