#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -x > gpurun_out/pytest42.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest42.log
grep -E "AssertionError|assert |passed|failed|Error" gpurun_out/pytest42.log | head -20
timeout 200 python tools/trace_step.py --layers 1 2>&1 | head -5
timeout 600 python bench.py --steps 64 --warmup 8 --no-extra > gpurun_out/bench42.json 2> gpurun_out/bench42.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench42.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac'])"
B200Q_FUSED=1 timeout 600 python bench.py --steps 64 --warmup 8 --no-extra > gpurun_out/bench42f.json 2> gpurun_out/bench42f.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench42f.json')); print('fused', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac'])"
timeout 200 python tools/prof_one.py --fmt Q4_K --N 28672 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 2 -c 1 -f -o gpurun_out/prof_matvec_q4k_gu_r1 python tools/prof_one.py --fmt Q4_K --N 28672 > gpurun_out/ncu_a.log 2>&1
tail -n 2 gpurun_out/ncu_a.log
