#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_quant.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for pf in 0 2 4; do echo "=== L2PF=$pf"; B200Q_MV_L2PF=$pf timeout 300 python tools/kbench.py --quick --fmts Q6_K 2>&1 | tail -4; done
