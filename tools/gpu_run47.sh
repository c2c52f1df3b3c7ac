#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --maxfail=5 -p no:cacheprovider 2>&1 | tail -5
timeout 600 python tools/kbench.py --fmts Q4_K,Q6_K,Q8_0 --ms 8,32,64,128 --quick 2>&1 | tail -48
timeout 600 python tools/kbench.py --fmts Q6_K --ms 2048 --quick 2>&1 | tail -4
