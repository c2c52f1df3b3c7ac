#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s); timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench exit $? wall $(( $(date +%s) - T0 )) s"
grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_default.err; grep -v "Warn\|^\s" gpurun_out/r2_bench_default.err | tail -5
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','n_gpus')}, 'e2e', j['e2e']['value'], 'launches', j['config']['launches_per_step'])
print('roofline', {k:j['roofline'][k] for k in ('achieved','frac','kernel','us_per_launch','step_frac')})
print('cpu', j.get('cpu_baseline',{}).get('value'), j.get('cpu_baseline',{}).get('cores'))
ex=j.get('extra',{})
for c in ex.get('configs',[]): print(c)
ms=ex.get('matvec_shapes')
if isinstance(ms,list):
    for r in ms: print(r)
else: print(ms)
for k in ex:
    if k not in ('configs','matvec_shapes'): print(k, ex[k])
PY
