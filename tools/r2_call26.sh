#!/bin/bash
# hoisted-address consumer loop + wide split-tile reducer: full GPU suite, then the default bench with extras
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/r2_pytest_gpu_3.tail
timeout 900 python bench.py > gpurun_out/r2_bench_c26.json 2> gpurun_out/r2_bench_c26.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_c26.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_bench_c26.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "ms_per_step")}, j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["us_per_launch"], j["roofline"]["step_frac"])
ex = j.get("extra", {})
for c in ex.get("configs", []): print(c)
for c in ex.get("matvec_shapes", []): print(c)
for k, v in ex.items():
    if k not in ("configs", "matvec_shapes"): print(k, v)
PY
