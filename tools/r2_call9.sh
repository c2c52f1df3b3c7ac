#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_step.py --workload mistral-7b:Q6_K --layers 2 > gpurun_out/r2_trace_step_7b_q6k.log 2>&1; cat gpurun_out/r2_trace_step_7b_q6k.log | grep -v Warn | head -12
timeout 300 python tools/trace_step.py --workload llama-3.2-1b:Q4_K_M --layers 2 > gpurun_out/r2_trace_step_1b.log 2>&1; cat gpurun_out/r2_trace_step_1b.log | grep -v Warn | head -10
