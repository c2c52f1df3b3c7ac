#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest5.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest5.log
tail -15 gpurun_out/pytest5.log
timeout 600 python tools/kbench.py --quick --json gpurun_out/kbench5.json > gpurun_out/kbench5.log 2>&1; echo "kbench exit $?" >> gpurun_out/kbench5.log
cat gpurun_out/kbench5.log
timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -6
