#!/bin/bash
timeout 120 python tools/trace_matvec.py --fmt Q6_K 2>&1 | tail -10
timeout 120 python tools/trace_matvec.py --fmt Q4_K --N 4096 --K 4096 2>&1 | tail -10
