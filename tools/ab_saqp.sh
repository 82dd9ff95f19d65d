#!/bin/bash
# A/B of the per-posterior Simpson kernel's CTA shape (threads, CTAs per SM, shared-memory KB
# per CTA) on C3 and on the resilience path.
B="python bench.py --steps 5 --warmup 3 --resilience-m 0 --gcp-candidates 0 --no-cpu-baseline"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print("C3 step", round(d["ms_per_step"],3), r["kernel"], round(r["kernel_ms_per_step"],3))'
for shape in "128 6 1" "128 6 10" "128 6 20" "256 3 1" "64 12 1" "192 4 1"; do
  set -- $shape
  echo "== threads $1 ctas $2 smem $3 KB"
  export RHFQ_SAQP_THREADS=$1 RHFQ_SAQP_CTAS=$2 RHFQ_SAQP_SMEM_KB=$3
  $B 2>/dev/null | python -c "$pick"
  for N in 50; do
    python tools/resil_trace.py $N 65520 4 2>&1 | grep "^run" | sort -k3 -n | head -1 | sed "s/^/resil N=$N best /"
  done
done
