#!/bin/bash
# TRACE build: per-launch timing structure of the 70B TP1 step (2 layers) and of rank 0's TP8 shard alone on one GPU
mkdir -p gpurun_out
timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp1.log; cat gpurun_out/r2_trace_step_70b_tp1.log | tail -14
B200Q_EMULATE_TP=8 timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp8emu_b.log; tail -14 gpurun_out/r2_trace_step_70b_tp8emu_b.log
for emu in 8; do
B200Q_EMULATE_TP=$emu timeout 300 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('emulate TP$emu (1 GPU, trace build):', {k:j.get(k) for k in ('value','ms_per_step')}, j['config'].get('launches_per_step'))
"
done
