#!/bin/bash
mkdir -p gpurun_out
for s in 0 1 2 3 4 8; do if [ $s = 0 ]; then timeout 200 python tools/gemm_m32_splits.py 2>&1 | grep SPLITS; else B200Q_GEMM_SPLITS=$s timeout 200 python tools/gemm_m32_splits.py 2>&1 | grep SPLITS; fi; done
