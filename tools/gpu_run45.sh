#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider 2>&1 | tail -5
timeout 600 python tools/kbench.py --fmts Q4_K,Q6_K,Q8_0 --ms 8,32,128 --quick 2>&1 | tail -36
timeout 600 python tools/kbench.py --fmts Q4_K,Q6_K --ms 2048 --quick 2>&1 | tail -8
