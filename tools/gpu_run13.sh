#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_quant.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest13.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest13.log
grep -E "AssertionError|assert |passed|failed|diverge" gpurun_out/pytest13.log | head -30
timeout 300 python tools/kbench.py --quick --fmts Q4_K,Q6_K,Q8_0 2>&1 | tail -13
