#!/bin/bash
# TP2: cluster consumer with one L2 round trip (epoch + both parities together), slot zeroing off the barrier path
mkdir -p gpurun_out
bash tools/r2_tp.sh 2 llama-3-70b:Q4_K_M 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -10
echo "== emulate TP8 on one GPU"
B200Q_EMULATE_TP=8 timeout 300 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | tail -1 | cut -c1-160
export B200Q_LIB=$PWD/blazr_b200/lib/libb200q_trace.so
B200Q_TRACE_DUMP=gpurun_out/r2_trace_tp8emu_d.npz B200Q_EMULATE_TP=8 timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 1 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp8emu_d.log; tail -9 gpurun_out/r2_trace_step_70b_tp8emu_d.log
