#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -x > gpurun_out/pytest41.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest41.log
grep -E "AssertionError|assert |passed|failed|Error" gpurun_out/pytest41.log | head -20
timeout 300 python tools/kbench.py --fmts Q4_K,Q6_K,Q8_0,AWQ --ms 1 --quick 2>&1 | tail -16
timeout 200 python tools/trace_step.py --layers 1 2>&1 | head -5
timeout 600 python bench.py --steps 64 --warmup 8 --no-extra > gpurun_out/bench41.json 2> gpurun_out/bench41.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench41.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac'])"
