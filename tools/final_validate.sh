#!/bin/bash
# One gpurun call that validates a build on ONE B200: full GPU suite, default bench with extras, filtered launch list of the default
# decode step, the reference (CPU) arm.  gpurun --timeout 2400 -- "bash tools/final_validate.sh"; multi-GPU: tools/r2_tp.sh N
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_final.tail
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "ms_per_step")}, j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["us_per_launch"], j["roofline"]["step_frac"], j.get("cpu_baseline", {}).get("value"))
ex = j.get("extra", {})
for c in ex.get("configs", []): print(c)
for k, v in ex.items():
    if k not in ("configs", "matvec_shapes"): print(k, str(v)[:300])
PY
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec_kernel|attn_decode|add_rmsnorm|argmax|embed_kernel|swiglu" -c 900 --csv --log-file gpurun_out/r2_decode_step_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc $?"; python tools/launch_share.py gpurun_out/r2_decode_step_launches.csv
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
