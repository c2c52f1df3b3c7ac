#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/tp_check.py 2>&1 | grep -v "^W\|Warning\|warn\|OMP_NUM\|^\*\*\*" | tail -6
for n in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 48 --warmup 8 2>gpurun_out/bench_tp$n.err | tail -1 > gpurun_out/bench_tp$n.json; tail -2 gpurun_out/bench_tp$n.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_tp$n.json')); print('TP$n', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 32 --warmup 8 --workload llama-3-70b:Q4_K_M 2>gpurun_out/bench_70b_tp8.err | tail -1 > gpurun_out/bench_70b_tp8.json; tail -2 gpurun_out/bench_70b_tp8.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/bench_70b_tp8.json')); print('70B TP8', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'])"
