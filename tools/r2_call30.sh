#!/bin/bash
# PDL on the tcgen05 path (stage / GEMM / split-K reduce): GEMM + batched decode parity, then the 7B bench with extras (batch-32, prefill)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_decode.py tests/test_gpu_parity_depth.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -5
timeout 600 python bench.py --workload mistral-7b:Q6_K > gpurun_out/r2_bench_c30.json 2> gpurun_out/r2_bench_c30.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_c30.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_bench_c30.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "ms_per_step")}, j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["us_per_launch"], j["roofline"]["step_frac"])
ex = j.get("extra", {})
for k, v in ex.items():
    if k not in ("configs", "matvec_shapes"): print(k, v)
PY
# kernel durations of the TP8-emulated step (serialised, cold): what the glue kernels cost
B200Q_EMULATE_TP=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec_kernel|attn_decode|add_rmsnorm|argmax|embed_kernel|swiglu" -c 700 --csv --log-file gpurun_out/r2_tp8emu_launches.csv python bench.py --steps 1 --warmup 3 --no-extra > gpurun_out/r2_ncu_tp8emu.log 2>&1
echo "rc $?"; python tools/launch_share.py gpurun_out/r2_tp8emu_launches.csv
