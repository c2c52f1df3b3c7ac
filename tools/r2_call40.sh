#!/bin/bash
# magic-number int->f64 in the matvec consumer loop (XU 7 -> 4 per chunk): parity subset + the 7B bench with the per-shape table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_gpu_decode.py tests/test_gpu_fused.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -3
B200Q_BENCH_EXTRA_S=120 timeout 600 python bench.py --workload mistral-7b:Q4_K > gpurun_out/r2_bench_c40.json 2> gpurun_out/r2_bench_c40.err; echo "bench exit $?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2_bench_c40.json").read().strip().splitlines()[-1])
print({k: j[k] for k in ("value", "ms_per_step")}, j["e2e"]["value"], j["roofline"]["frac"], j["roofline"]["us_per_launch"], j["roofline"]["step_frac"])
ex = j.get("extra", {})
for c in ex.get("configs", []): print(c)
for c in ex.get("matvec_shapes", []): print(c)
PY
