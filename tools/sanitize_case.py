"""Small end-to-end case for compute-sanitizer (posterior, resilience, GCP)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from reheatfunq_b200 import PosteriorBatch
from reheatfunq_b200.resilience import resilience_run
from reheatfunq_b200.regional import backend as gb

q, c, off = bench.make_regions(24, seed=3)
B = PosteriorBatch.from_packed(q, c, off, np.ones(24), np.arange(25, dtype=np.int64), bench.DEFAULT_PRIOR,
                               bli_max_refinements=5, rtol=1e-6)
x = np.linspace(0.05, 0.95, 5)[None, :] * B.Qmax()[:, None]
print(B.pdf(x).sum(), B.cdf(x).sum(), B.tail_quantiles(np.array([0.1, 0.5, 0.9])).sum(), B.status().sum())
for alg in ("explicit", "adaptive_simpson"):
    B2 = PosteriorBatch.from_packed(q[:off[3]], c[:off[3]], off[:4], np.ones(3), np.arange(4, dtype=np.int64),
                                    bench.DEFAULT_PRIOR, pdf_algorithm=alg)
    print(alg, B2.pdf(x[:3]).sum(), B2.cdf(x[:3]).sum())
res, fail, st = resilience_run(1, (31, 3.5, 0.33, 62, 5.5, 0.67), 20, 64, 100.0, np.linspace(0.9, 0.1, 5),
                               bench.DEFAULT_PRIOR, 1.0, 1e-4, 7, nstreams=2, chunk=24)
print(np.nansum(res), fail)
params = np.array([[300.0, 5000.0, 70.0, 70.0], [70.0, 1180.0, 17.0, 17.0]])
print(gb.gamma_conjugate_prior_kullback_leibler_batch_max_many(params, [[1.0, 10.0, 1.5, 1.0], [0.9, 15.4, 0.22, 0.218]]))
