#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tp.py tests/test_gpu_decode.py -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -15
for wl in mistral-7b:Q6_K llama-3.2-1b:Q4_K_M; do
for old in 1 0; do
  echo "== $wl NORM_OLD=$old"
  B200Q_NORM_OLD=$old timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-extra 2>gpurun_out/r2_c5.err | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step')}, j['e2e']['value'], j['config'].get('launches_per_step'), j['roofline']['step_frac'], j['roofline']['frac'])
"
  tail -2 gpurun_out/r2_c5.err
done
done
