#!/bin/bash
# first GPU session: parity tests, kernel bench, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/pytest1.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest1.log
tail -30 gpurun_out/pytest1.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke1.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke1.log; tail -5 gpurun_out/smoke1.log
timeout 600 python tools/kbench.py --quick --json gpurun_out/kbench1.json > gpurun_out/kbench1.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench1.log
cat gpurun_out/kbench1.log
timeout 200 python tools/prof_one.py --fmt Q6_K > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tools/prof_one.py --fmt Q6_K > gpurun_out/ncu1.log 2>&1
timeout 200 python tools/prof_one.py --fmt Q6_K > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 2 -c 2 -f -o gpurun_out/prof_q6k_r1 python tools/prof_one.py --fmt Q6_K > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out
