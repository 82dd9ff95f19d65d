#!/bin/bash
# A/B of alternative builds (build/variants/librhfq_NAME.so, tools/build_variant.sh) on C3 and
# on the resilience path.  usage: tools/ab_lib.sh NAME...   ("default" = in-tree library)
B="python bench.py --steps 5 --warmup 3 --resilience-m 0 --gcp-candidates 0 --no-cpu-baseline"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; o=d["roofline_other"]; print("C3 step", round(d["ms_per_step"],3), r["kernel"], round(r["kernel_ms_per_step"],3), o["kernel"], round(o["kernel_ms_per_step"],3))'
for v in "$@"; do
  if [ "$v" = default ]; then unset RHFQ_LIB; else export RHFQ_LIB=/root/repo/build/variants/librhfq_$v.so; fi
  echo "== $v"
  $B 2>/dev/null | python -c "$pick"
  python tools/resil_trace.py 50 65520 4 2>&1 | grep "^run" | sort -k3 -n | head -1 | sed "s/^/resil N=50 best /"
done
