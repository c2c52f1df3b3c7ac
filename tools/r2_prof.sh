#!/bin/bash
# round-2 evidence: launch list of the default bench step + ncu --set full captures of the roofline kernel and the prefill GEMM
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec_kernel|attn_decode|add_rmsnorm|argmax|embed_kernel|swiglu" -c 900 --csv --log-file gpurun_out/r2_decode_step_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc $?"; wc -l gpurun_out/r2_decode_step_launches.csv
timeout 300 python tools/prof_gu.py > gpurun_out/r2_prof_gu_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_matvec_q4k_gu python tools/prof_gu.py > gpurun_out/r2_ncu_gu.log 2>&1
echo "gu capture rc $?"; tail -2 gpurun_out/r2_ncu_gu.log
timeout 300 python tools/prof_gu.py --fmt Q6_K --F 14336 --K 4096 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_matvec_q6k_gu python tools/prof_gu.py --fmt Q6_K --F 14336 --K 4096 > gpurun_out/r2_ncu_gu6.log 2>&1
echo "gu q6k capture rc $?"
timeout 300 python tools/prof_one.py --fmt Q6_K --N 14336 --K 4096 --M 2048 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r2_prof_gemm_q6k_prefill python tools/prof_one.py --fmt Q6_K --N 14336 --K 4096 --M 2048 > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm capture rc $?"
ls -la gpurun_out/*.ncu-rep
