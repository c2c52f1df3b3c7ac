#!/bin/bash
# dual-format q|k + v launch: parity, then the default bench and the TP8-shard bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dual.py tests/test_gpu_decode.py tests/test_gpu_fused.py tests/test_gpu_parity_depth.py tests/test_gpu_tp.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -12
timeout 600 python bench.py --no-extra 2>/dev/null | tail -1 | cut -c1-200
B200Q_DUAL=0 timeout 600 python bench.py --no-extra 2>/dev/null | tail -1 | cut -c1-200
B200Q_EMULATE_TP=8 timeout 300 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | tail -1 | cut -c1-200
