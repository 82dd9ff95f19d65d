#!/bin/bash
# A/B of alternative builds of the CUDA library (tools/build_variant.sh): per-stage device times
# of the bench step.  usage: tools/ab_variants.sh [variant names...]   ("" = in-tree default)
cd /root/repo
for v in "" "$@"; do
  lib=${v:+/root/repo/build/variants/librhfq_$v.so}
  echo "== lib=${v:-default}"
  RHFQ_LIB=$lib python - <<'PY'
import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, bench, torch
from reheatfunq_b200 import PosteriorBatch
q,c,off = bench.make_regions(10000, seed=1)
for it in range(5):
    torch.cuda.synchronize(); t0=time.perf_counter()
    B = PosteriorBatch.from_packed(q,c,off,np.ones(10000),np.arange(10001,dtype=np.int64),bench.DEFAULT_PRIOR)
    B.tail_quantiles(bench.QUANTILES); torch.cuda.synchronize(); t1=time.perf_counter()
    cnt=B.counters()
print("step %.2f ms  node %.2f ms norm %.2f ms plan %.2f ms saq %.2f ms  leaves %d" % ((t1-t0)*1e3, cnt['node_kernel_ns']*1e-6, cnt['norm_kernel_ns']*1e-6, cnt['plan_kernel_ns']*1e-6, cnt['saq_kernel_ns']*1e-6, B.saq_leaves()))
PY
done
