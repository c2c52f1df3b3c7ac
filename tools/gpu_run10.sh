#!/bin/bash
echo "=== Q6_K"; timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -4
echo "=== Q6_K nocompute"; B200Q_MV_DEBUG=1 timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -4
