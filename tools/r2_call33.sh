#!/bin/bash
# TRACE build: matvec + glue kernel timeline of rank 0's TP8 shard (one GPU) and of the TP1 step
mkdir -p gpurun_out
B200Q_EMULATE_TP=8 timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp8emu_c.log; tail -26 gpurun_out/r2_trace_step_70b_tp8emu_c.log
timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 1 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp1_c.log; tail -9 gpurun_out/r2_trace_step_70b_tp1_c.log
