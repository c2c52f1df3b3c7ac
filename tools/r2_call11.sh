#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -p no:cacheprovider --timeout 300 2>&1 | tail -6
run() { # label, env...
  label=$1; shift
  for wl in mistral-7b:Q6_K llama-3.2-1b:Q4_K_M mistral-7b:Q4_K; do
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
