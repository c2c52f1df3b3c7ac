#!/bin/bash
# multi-GPU validation: gpurun --gpus N -- 'bash tools/r2_tp.sh N [workloads...]'
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_tp.py -m gpu -q -p no:cacheprovider --timeout 600 -k "tensor_parallel" 2>&1 | tail -6
cat gpurun_out/tp_check_world$N.log | grep "^tp\|TP_CHECK" 
for wl in ${@:-llama-3-70b:Q4_K_M}; do
  echo "== $wl TP$N"
  B200Q_BENCH_WATCHDOG_S=500 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $wl --steps 64 --warmup 8 --no-extra 2>gpurun_out/r2_bench_tp$N.err | tail -1 > gpurun_out/r2_bench_tp${N}_${wl%%:*}.json
  tail -3 gpurun_out/r2_bench_tp$N.err | cut -c1-300; python -c "
import json,sys
j=json.load(open('gpurun_out/r2_bench_tp${N}_${wl%%:*}.json'))
print({k:j.get(k) for k in ('value','ms_per_step','n_gpus')}, j['e2e']['value'], j['config'].get('launches_per_step'), j['roofline']['step_frac'], j['roofline']['frac'])
"
done
