#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_depth.py -m gpu -q -p no:cacheprovider --timeout 900 -k "32_layers or prefill" --tb=short 2>&1 | grep -v "Warning\|warn" | tail -60
