#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest24.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest24.log
grep -E "AssertionError|assert |passed|failed|Error" gpurun_out/pytest24.log | head -20
timeout 900 python bench.py > gpurun_out/bench24.json 2> gpurun_out/bench24.err; echo "bench exit $?"
tail -2 gpurun_out/bench24.err; python -c "
import json; d=json.load(open('gpurun_out/bench24.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['step_frac'], d.get('extra'), d['clocks'])"
timeout 200 python tools/prof_one.py --fmt Q6_K --N 28672 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 2 -c 2 -f -o gpurun_out/prof_matvec_q6k_gu_r1 python tools/prof_one.py --fmt Q6_K --N 28672 > gpurun_out/ncu_a.log 2>&1
tail -n 2 gpurun_out/ncu_a.log
timeout 200 python tools/prof_one.py --fmt Q6_K --N 14336 --M 2048 --iters 4 > gpurun_out/plain_g.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 2 -f -o gpurun_out/prof_gemm_q6k_r1 python tools/prof_one.py --fmt Q6_K --N 14336 --M 2048 --iters 4 > gpurun_out/ncu_b.log 2>&1
tail -n 2 gpurun_out/ncu_b.log
