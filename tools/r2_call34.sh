#!/bin/bash
mkdir -p gpurun_out
B200Q_TRACE_DUMP=gpurun_out/r2_trace_tp8emu.npz B200Q_EMULATE_TP=8 timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp8emu_c.log; tail -3 gpurun_out/r2_trace_step_70b_tp8emu_c.log
B200Q_TRACE_DUMP=gpurun_out/r2_trace_tp1.npz timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp1_c.log; tail -3 gpurun_out/r2_trace_step_70b_tp1_c.log
B200Q_TRACE_DUMP=gpurun_out/r2_trace_7b.npz timeout 300 python tools/trace_step.py --workload mistral-7b:Q4_K --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_7b.log; tail -3 gpurun_out/r2_trace_step_7b.log
B200Q_TRACE_DUMP=gpurun_out/r2_trace_1b.npz timeout 300 python tools/trace_step.py --workload llama-3.2-1b:Q4_K_M --layers 2 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_1b.log; tail -3 gpurun_out/r2_trace_step_1b.log
