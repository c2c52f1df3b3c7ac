#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_decode.py tests/test_gpu_tp.py tests/test_gpu_moe.py -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -8
timeout 300 python tools/trace_step.py --workload mistral-7b:Q6_K --layers 2 > gpurun_out/r2_trace_step_7b_q6k_fused2.log 2>&1; grep -v Warn gpurun_out/r2_trace_step_7b_q6k_fused2.log | head -10
