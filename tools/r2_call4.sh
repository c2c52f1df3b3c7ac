#!/bin/bash
# successor L2 prefetch (b200q_weight_set_next): kernel chain with / without, decode step with / without, budgets
mkdir -p gpurun_out
{
echo "### kbench quick, no chain"; timeout 300 python tools/kbench.py --quick --fmts Q4_K,Q8_0 2>&1 | grep -v Warn
for mb in 16 32 64; do
echo "### kbench quick, chain, PF_MB=$mb"; B200Q_MV_NEXT_PF_MB=$mb timeout 300 python tools/kbench.py --quick --fmts Q4_K,Q8_0 --chain 2>&1 | grep -v Warn
done
echo "### trace chain Q4_K 28672x4096"; timeout 200 python tools/trace_matvec.py --fmt Q4_K --N 28672 --K 4096 --n 5 --chain 2>&1 | tail -6
for wl in mistral-7b:Q4_K mistral-7b:Q6_K llama-3.2-1b:Q4_K_M; do
for ch in 0 1; do
  echo "== $wl CHAIN=$ch"
  B200Q_CHAIN=$ch timeout 300 python bench.py --workload $wl --steps 64 --warmup 8 --no-extra 2>gpurun_out/r2_chain_$ch.err | python -c "
import json,sys
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print({k:j.get(k) for k in ('value','ms_per_step')}, j['e2e']['value'], j['config'].get('launches_per_step'), j['roofline']['step_frac'], j['roofline']['frac'])
"
done
done
} > gpurun_out/r2_chain_1.log 2>&1
cat gpurun_out/r2_chain_1.log
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_quant.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
