"""Device times of the timed kernels for one bench step (CUDA-event timers inside the library)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from reheatfunq_b200 import PosteriorBatch

R = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
q, c, off = bench.make_regions(R, seed=1)
for it in range(2):
    B = PosteriorBatch.from_packed(q, c, off, np.ones(R), np.arange(R + 1, dtype=np.int64),
                                   bench.DEFAULT_PRIOR)
    B.tail_quantiles(bench.QUANTILES)
    cnt = B.counters()
    del B
print("node plan %.2f ms | node quad %.2f ms | norm quad %.2f ms | saq step %.2f ms | g_evals %d (quad %d) newton %d"
      % (cnt["plan_kernel_ns"] * 1e-6, cnt["node_kernel_ns"] * 1e-6, cnt["norm_kernel_ns"] * 1e-6,
         cnt["saq_kernel_ns"] * 1e-6, cnt["g_evals"], cnt["g_evals_quad"], cnt["newton"]))
