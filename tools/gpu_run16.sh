#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest16.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest16.log
grep -E "AssertionError|assert |passed|failed|diverge|Error" gpurun_out/pytest16.log | head -20
timeout 900 python bench.py --steps 64 --warmup 8 --no-extra > gpurun_out/bench16.json 2> gpurun_out/bench16.err; echo "bench exit $?"
tail -3 gpurun_out/bench16.err; python -c "
import json; d=json.load(open('gpurun_out/bench16.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac'])"
bash tools/gpu_run15.sh 2>&1 | tail -9
