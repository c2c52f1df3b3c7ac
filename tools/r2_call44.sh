#!/bin/bash
# validated state of the current build: full GPU suite + default bench (no extras)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_final.tail
timeout 600 python bench.py --no-extra 2>/dev/null | tail -1 | cut -c1-330
