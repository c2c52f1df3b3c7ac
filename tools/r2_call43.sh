#!/bin/bash
# 4 GPUs, final build: MoE / quant parity on GPU 0, then tensor-parallel parity (tp_check world 4) + 70B Q4_K_M TP4 bench + EP4
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_moe.py tests/test_gpu_quant.py tests/test_gpu_tp.py -m gpu -q -p no:cacheprovider --timeout 300 -k "not tensor_parallel" 2>&1 | tail -3
bash tools/r2_tp.sh 4 llama-3-70b:Q4_K_M 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 tools/ep_check.py 2>&1 | grep -v "^W\|Warning\|warn\|OMP_NUM\|^\*\*\*" | tail -3 | tee gpurun_out/ep_check_world4.log
