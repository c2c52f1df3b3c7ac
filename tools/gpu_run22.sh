#!/bin/bash
for pf in 0 8 24 48; do echo "=== L2 prefetch $pf MB"; B200Q_PF_MB=$pf timeout 600 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'gu frac', round(d['roofline']['frac'],3), 'step frac', round(d['roofline']['step_frac'],3))"; done
