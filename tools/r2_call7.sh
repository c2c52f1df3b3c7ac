#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -4
run() { # label, dir, env...
  label=$1; dir=$2; shift 2
  for wl in mistral-7b:Q6_K llama-3.2-1b:Q4_K_M; do
  echo "== $label $wl"
  (cd $dir && env "$@" timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step')}, j['e2e']['value'], j['config'].get('launches_per_step'), j['roofline']['step_frac'], j['roofline']['frac'])
")
  done
}
run old _old X=1
run new-default . X=1
run new-chain0 . B200Q_CHAIN=0
