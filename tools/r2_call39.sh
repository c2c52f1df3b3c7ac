#!/bin/bash
# final-build validation: full GPU suite, default bench with extras, launch list + ncu captures (each only after its command exited 0 without ncu)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_pytest_gpu_final.tail
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "ms_per_step")}, j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["us_per_launch"], j["roofline"]["step_frac"], j.get("cpu_baseline", {}).get("value"))
ex = j.get("extra", {})
for c in ex.get("configs", []): print(c)
for c in ex.get("matvec_shapes", []): print(c)
for k, v in ex.items():
    if k not in ("configs", "matvec_shapes"): print(k, v)
PY
bash tools/r2_prof.sh 2>&1 | tail -12
