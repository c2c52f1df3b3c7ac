#!/bin/bash
# TP8 latency structure on ONE GPU: rank 0's shard of the 70B model with the world-1 fused exchange (TRACE build of the library)
mkdir -p gpurun_out
B200Q_EMULATE_TP=8 timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 2 > gpurun_out/r2_trace_step_70b_tp8emu.log 2>&1; grep -v Warn gpurun_out/r2_trace_step_70b_tp8emu.log | tail -14
for emu in 8 4 2; do
B200Q_EMULATE_TP=$emu timeout 300 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('emulate TP$emu (1 GPU, trace build):', {k:j.get(k) for k in ('value','ms_per_step')}, j['config'].get('launches_per_step'))
"
done
