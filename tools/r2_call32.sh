#!/bin/bash
# 8 GPUs: tensor-parallel parity (tp_check world 8 through tests/test_gpu_tp.py), 70B Q4_K_M TP8 bench, expert-parallel check
mkdir -p gpurun_out
bash tools/r2_tp.sh 8 llama-3-70b:Q4_K_M 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -9
echo "== expert parallel, 8 GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 tools/ep_check.py 2>&1 | grep -v "^W\|Warning\|warn\|OMP_NUM\|^\*\*\*" | tail -4 | tee gpurun_out/ep_check_world8.log
