#!/bin/bash
timeout 600 python tools/kbench.py --fmts Q6_K --ms 32,2048 --quick 2>&1 | tail -8
