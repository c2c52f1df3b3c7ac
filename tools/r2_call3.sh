#!/bin/bash
mkdir -p gpurun_out
for spec in "Q4_K 28672 4096" "Q4_K 4096 4096" "Q8_0 28672 4096" "Q4_K 4096 14336"; do
  set -- $spec
  timeout 200 python tools/trace_matvec.py --fmt $1 --N $2 --K $3 --n 6 2>&1 | grep -v Warn | tail -8
done > gpurun_out/r2_trace_matvec_1.log 2>&1
cat gpurun_out/r2_trace_matvec_1.log
