#!/bin/bash
# A/B of environment switches on the C3 step: tools/ab_env.sh "VAR=1 VAR2=x" "..." ("-" = none)
B="python bench.py --steps 5 --warmup 3 --resilience-m 0 --gcp-candidates 0 --no-cpu-baseline"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; o=d["roofline_other"]; print("C3 step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), r["kernel"], round(r["kernel_ms_per_step"],3), o["kernel"], round(o["kernel_ms_per_step"],3))'
for v in "$@"; do
  echo "== $v"
  if [ "$v" = "-" ]; then $B 2>/dev/null | python -c "$pick"; else env $v $B 2>/dev/null | python -c "$pick"; fi
done
