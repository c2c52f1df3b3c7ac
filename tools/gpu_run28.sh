#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_quant.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest28.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest28.log
tail -3 gpurun_out/pytest28.log
B200Q_FUSED=1 timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
for f in 1 0; do echo "=== fused norm=$f"; B200Q_FUSED=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:round(d[k],2) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'gu frac', round(d['roofline']['frac'],3), 'step frac', round(d['roofline']['step_frac'],3), d['config']['launches_per_step'])"; done
