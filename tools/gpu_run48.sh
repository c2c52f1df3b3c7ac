#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_gpu_decode.py -m gpu -q -k "peer_allreduce" -p no:cacheprovider 2>&1 | tail -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py 2>&1 | grep -v "^W\|Warning\|warn" | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 64 --warmup 8 2>gpurun_out/bench_tp2.err | tail -1 > gpurun_out/bench_tp2.json; tail -3 gpurun_out/bench_tp2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tp2.json')); print('TP2 peer', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'])"
B200Q_TP_NCCL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 64 --warmup 8 2>gpurun_out/bench_tp2n.err | tail -1 > gpurun_out/bench_tp2n.json; tail -3 gpurun_out/bench_tp2n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_tp2n.json')); print('TP2 nccl', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['roofline']['frac'])"
