#!/bin/bash
# TP2: element-wise exchange consumer + fused prologue matvecs (B200Q_TP_FINISH=1) vs the cluster consumer (=0)
mkdir -p gpurun_out
bash tools/r2_tp.sh 2 llama-3-70b:Q4_K_M 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -9
echo "== cluster consumer (B200Q_TP_FINISH=0)"
B200Q_TP_FINISH=0 B200Q_BENCH_WATCHDOG_S=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 64 --warmup 8 --no-extra 2>/dev/null | tail -1 | cut -c1-160
for f in 1 0; do
echo "== emulate TP8 on one GPU, B200Q_TP_FINISH=$f"
B200Q_TP_FINISH=$f B200Q_EMULATE_TP=8 timeout 300 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | tail -1 | cut -c1-160
done
