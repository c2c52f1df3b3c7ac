#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_moe.py -m gpu -q --maxfail=5 -p no:cacheprovider 2>&1 | tail -15
timeout 200 python tools/prof_one.py --fmt Q8_0 --N 14336 --M 32 --iters 4 > gpurun_out/plain_g.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_gemm_q80_m32_r1b python tools/prof_one.py --fmt Q8_0 --N 14336 --M 32 --iters 4 > gpurun_out/ncu_b.log 2>&1
tail -n 2 gpurun_out/ncu_b.log
