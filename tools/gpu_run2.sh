#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest2.log
tail -5 gpurun_out/pytest2.log
timeout 600 python tools/kbench.py --quick --json gpurun_out/kbench2.json > gpurun_out/kbench2.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench2.log
cat gpurun_out/kbench2.log
timeout 200 python tools/prof_one.py --fmt Q6_K > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 2 -c 2 -f -o gpurun_out/prof_q6k_r2 python tools/prof_one.py --fmt Q6_K > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/ncu2.log
