#!/bin/bash
# ncu --set full captures of the shipped kernels (final build; each only after its command exited 0 without ncu)
mkdir -p gpurun_out
timeout 300 python tools/prof_gu.py > gpurun_out/r2_prof_gu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_matvec_q4k_gu python tools/prof_gu.py > gpurun_out/r2_ncu_gu.log 2>&1
echo "gu capture rc $?"
timeout 300 python tools/prof_gu.py --fmt Q6_K --F 14336 --K 4096 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_matvec_q6k_gu python tools/prof_gu.py --fmt Q6_K --F 14336 --K 4096 > gpurun_out/r2_ncu_gu6.log 2>&1
echo "gu q6k capture rc $?"
timeout 300 python tools/prof_one.py --fmt Q6_K --N 14336 --K 4096 --M 2048 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_gemm_q6k_prefill python tools/prof_one.py --fmt Q6_K --N 14336 --K 4096 --M 2048 > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm capture rc $?"
