#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider 2>&1 | tail -8
timeout 900 python bench.py --steps 64 --warmup 8 --no-extra > gpurun_out/bench50.json 2> gpurun_out/bench50.err; echo "bench exit $?"; tail -3 gpurun_out/bench50.err
python -c "
import json; d=json.load(open('gpurun_out/bench50.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['us_per_launch'], d['roofline']['step_frac'])"
