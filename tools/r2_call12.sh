#!/bin/bash
mkdir -p gpurun_out
B200Q_FUSED=1 timeout 300 python tools/trace_step.py --workload mistral-7b:Q6_K --layers 2 > gpurun_out/r2_trace_step_7b_q6k_fused.log 2>&1; grep -v Warn gpurun_out/r2_trace_step_7b_q6k_fused.log | head -10
