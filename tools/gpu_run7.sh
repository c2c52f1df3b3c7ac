#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest7.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest7.log
tail -8 gpurun_out/pytest7.log
timeout 600 python tools/kbench.py --quick --json gpurun_out/kbench7.json > gpurun_out/kbench7.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench7.log
cat gpurun_out/kbench7.log
timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -7
echo "--- no L2 prefetch"; B200Q_MV_L2PF=0 timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -4
B200Q_MV_L2PF=0 timeout 300 python tools/kbench.py --quick --fmts Q6_K 2>&1 | tail -5
