#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -rxXf 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu_2.tail; cat gpurun_out/r2_pytest_gpu_2.tail
