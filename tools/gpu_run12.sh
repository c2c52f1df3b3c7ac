#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_decode.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_decode.log
tail -40 gpurun_out/pytest_decode.log
