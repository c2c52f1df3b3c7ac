#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q --maxfail=10 -p no:cacheprovider -x > gpurun_out/pytest43.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest43.log
grep -E "AssertionError|assert |passed|failed|Error" gpurun_out/pytest43.log | head -20
timeout 200 python tools/trace_step.py --layers 1 2>&1 | head -5
timeout 600 python bench.py --steps 64 --warmup 8 --no-extra > gpurun_out/bench43.json 2> gpurun_out/bench43.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench43.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_frac'])"
timeout 200 python tools/prof_one.py --fmt Q6_K --N 14336 --M 32 --iters 4 > gpurun_out/plain_g.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_gemm_q6k_m32_r1 python tools/prof_one.py --fmt Q6_K --N 14336 --M 32 --iters 4 > gpurun_out/ncu_b.log 2>&1
tail -n 2 gpurun_out/ncu_b.log
