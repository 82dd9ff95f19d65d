for lib in "" build/variants/librhfq_mb5.so build/variants/librhfq_mb6.so; do
  echo "== lib=${lib:-default}"
  RHFQ_LIB=${lib:+/root/repo/$lib} python tools/profile_step.py 10000 3 2>&1 | tail -1
  RHFQ_LIB=${lib:+/root/repo/$lib} python - <<'PY'
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, bench
from reheatfunq_b200 import PosteriorBatch
q,c,off = bench.make_regions(10000, seed=1)
B = PosteriorBatch.from_packed(q,c,off,np.ones(10000),np.arange(10001,dtype=np.int64),bench.DEFAULT_PRIOR)
B.tail_quantiles(bench.QUANTILES); cnt=B.counters()
print("node %.2f ms norm %.2f ms saq %.2f ms" % (cnt['node_kernel_ns']*1e-6, cnt['norm_kernel_ns']*1e-6, cnt['saq_kernel_ns']*1e-6))
PY
done
