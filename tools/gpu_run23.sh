#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py 2>&1 | grep -v "^W\|Warning\|warn" | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 64 --warmup 8 2>gpurun_out/bench_tp2.err | tail -1 > gpurun_out/bench_tp2.json; tail -3 gpurun_out/bench_tp2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tp2.json')); print('TP2', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'])"
