#!/bin/bash
# round-2 GPU call 2: persistent op-list kernel (B200Q_DSTEP) vs launch-per-op, whole-step tok/s
mkdir -p gpurun_out
for wl in mistral-7b:Q6_K mistral-7b:Q4_K llama-3.2-1b:Q4_K_M; do
for d in 0 1 2; do
  echo "== $wl DSTEP=$d"
  B200Q_DSTEP=$d timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-extra 2>gpurun_out/r2_dstep_$d.err | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step')}, j['e2e']['value'], j['config'].get('launches_per_step'), j['roofline']['step_frac'])
"
done
done
