#!/bin/bash
# decode tok/s of every BASELINE.json config that fits one B200 (batch 1, graph replay), one JSON line each
mkdir -p gpurun_out
out=gpurun_out/config_sweep.jsonl
: > $out
for wl in llama-3.2-1b:Q4_K_M mistral-7b:Q6_K mistral-7b:Q8_0 mistral-7b:Q4_K llama-3-8b:AWQ llama-3-8b:GPTQ llama-3-70b:Q4_K_M; do
  timeout 900 python bench.py --workload $wl --steps 48 --warmup 8 --no-extra 2> gpurun_out/sweep_err.log | tail -1 >> $out || echo "{\"workload\": \"$wl\", \"error\": \"failed\"}" >> $out
done
python - <<'PY'
import json
for line in open('gpurun_out/config_sweep.jsonl'):
    try:
        d = json.loads(line)
        r = d['roofline']
        print(f"{d['config']['workload'][:44]:44s} {d['value']:8.1f} tok/s  e2e {d['e2e']['value']:8.1f}  step {d['ms_per_step']:7.3f} ms  step_frac_hbm {r['step_frac']:.3f}  gate|up {r['us_per_launch']:6.2f} us frac {r['frac']:.3f}  cpu {d.get('cpu_baseline',{}).get('value')}")
    except Exception as e:
        print('ERR', line[:200], e)
PY
