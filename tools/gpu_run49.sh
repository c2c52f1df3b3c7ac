#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider 2>&1 | tail -8
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench49.json 2> gpurun_out/bench49.err; echo "bench exit $?"; tail -3 gpurun_out/bench49.err
python -c "
import json; d=json.load(open('gpurun_out/bench49.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['step_frac']); print(json.dumps(d.get('extra'), indent=1))"
