#!/bin/bash
# TP2: xk_off fix + self-validating all-reduce slots: fused / depth / tp tests, tp_check world 2, 70B TP2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity_depth.py tests/test_gpu_tp.py tests/test_gpu_moe.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -6
bash tools/r2_tp.sh 2 llama-3-70b:Q4_K_M 2>&1 | tail -12
