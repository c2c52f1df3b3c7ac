#!/bin/bash
echo "=== Q8_0 14336x4096"; timeout 120 python tools/trace_matvec.py --fmt Q8_0 2>&1 | tail -5
echo "=== Q8_0 nocompute"; B200Q_MV_DEBUG=1 timeout 120 python tools/trace_matvec.py --fmt Q8_0 2>&1 | tail -5
echo "=== Q6_K nocompute"; B200Q_MV_DEBUG=1 timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -5
echo "=== Q6_K 28672"; timeout 120 python tools/trace_matvec.py --fmt Q6_K --N 28672 2>&1 | tail -5
echo "=== Q6_K grid 296"; B200Q_MV_GRID=296 timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -5
