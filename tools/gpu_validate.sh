#!/bin/bash
# One gpurun call that re-validates a build on a B200: the GPU parity suite, the default bench line, the launch list of
# the same command and one `ncu --set full` capture of the roofline kernel (each ncu pass only after its command has
# exited 0 without ncu).  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'
# Multi-GPU (N = 2, 4, 8 -- charged N x): gpurun --gpus N -- 'bash tools/gpu_validate.sh tp N'
mkdir -p gpurun_out
if [ "$1" = "tp" ]; then
  N=${2:-2}
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py 2>&1 | grep -v "^W\|Warning\|warn\|OMP_NUM\|^\*\*\*" | tail -8
  B200Q_BENCH_WATCHDOG_S=240 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 64 --warmup 8 2>gpurun_out/bench_tp$N.err | tail -1 > gpurun_out/bench_tp$N.json
  tail -2 gpurun_out/bench_tp$N.err | cut -c1-300; cut -c1-400 gpurun_out/bench_tp$N.json
  exit 0
fi
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider -rX 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_launches.log 2>&1
timeout 200 python tools/prof_one.py --fmt Q6_K --N 28672 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matvec_kernel -s 2 -c 1 -f -o gpurun_out/prof_matvec_q6k_gu python tools/prof_one.py --fmt Q6_K --N 28672 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
