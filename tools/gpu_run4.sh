#!/bin/bash
mkdir -p gpurun_out
run() { echo "=== $*"; env "$@" timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -6; }
run B200Q_MV_DEBUG=1
run B200Q_MV_DEBUG=1 B200Q_MV_STAGES=4
run B200Q_MV_DEBUG=1 B200Q_MV_STAGES=2
run B200Q_MV_DEBUG=0 B200Q_MV_STAGES=4
echo "=== Q8_0 nocompute"; B200Q_MV_DEBUG=1 timeout 120 python tools/trace_matvec.py --fmt Q8_0 --N 28672 2>&1 | tail -6
echo "=== Q8_0 compute"; timeout 120 python tools/trace_matvec.py --fmt Q8_0 --N 28672 2>&1 | tail -6
