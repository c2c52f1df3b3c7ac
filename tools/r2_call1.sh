#!/bin/bash
# round-2 GPU call 1: the whole -m gpu suite (no -x), experimental dstep under timeout, per-shape kernel table, 70B N=1 baseline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -rxXfE --timeout 300 > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/r2_pytest_gpu_1.log
B200Q_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_zz_gpu_new_formats.py -m gpu -q -p no:cacheprovider -rxXfE -k dstep --timeout 120 > gpurun_out/r2_pytest_dstep_1.log 2>&1; echo "dstep exit $?"
tail -5 gpurun_out/r2_pytest_dstep_1.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 600 python tools/kbench.py --quick --json gpurun_out/r2_kbench_1.json > gpurun_out/r2_kbench_1.log 2>&1; echo "kbench exit $?"; cat gpurun_out/r2_kbench_1.log | tail -20
timeout 900 python bench.py --workload llama-3-70b:Q4_K_M --steps 32 --warmup 4 --no-extra > gpurun_out/r2_bench70b_1.json 2> gpurun_out/r2_bench70b_1.err; echo "bench70b exit $?"; cut -c1-600 gpurun_out/r2_bench70b_1.json
