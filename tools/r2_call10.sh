#!/bin/bash
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  for wl in mistral-7b:Q6_K llama-3.2-1b:Q4_K_M; do
  echo "== $label $wl"
  env "$@" timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step')}, j['e2e']['value'], j['config'].get('launches_per_step'), j['roofline']['step_frac'], j['roofline']['frac'])
"
  done
}
run default X=1
run fused-norm B200Q_FUSED=1
run fused-norm-swiglu B200Q_FUSED=1 B200Q_FUSED_SWIGLU=1
