#!/bin/bash
mkdir -p gpurun_out
for f in Q6_K Q4_K Q8_0; do timeout 120 python tools/trace_matvec.py --fmt $f > gpurun_out/trace_$f.log 2>&1; cat gpurun_out/trace_$f.log; done
B200Q_MV_STAGES=3 timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -7
timeout 120 python tools/trace_matvec.py --fmt Q4_K --N 4096 --K 4096 2>&1 | tail -7
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q --maxfail=5 -p no:cacheprovider -x > gpurun_out/pytest_gemm.log 2>&1
echo "pytest gemm exit $?" >> gpurun_out/pytest_gemm.log
tail -40 gpurun_out/pytest_gemm.log
