#!/bin/bash
mkdir -p gpurun_out
export B200Q_LIB=$PWD/blazr_b200/lib/libb200q_trace.so
B200Q_TRACE_DUMP=gpurun_out/r2_trace_tp8emu_e.npz B200Q_EMULATE_TP=8 timeout 300 python tools/trace_step.py --workload llama-3-70b:Q4_K_M --layers 1 2>&1 | grep -v Warn > gpurun_out/r2_trace_step_70b_tp8emu_e.log; tail -3 gpurun_out/r2_trace_step_70b_tp8emu_e.log
python - <<'PY'
import numpy as np
z = np.load("gpurun_out/r2_trace_tp8emu_e.npz"); gl = z["glue"]; gl = gl[gl[:, 1] > 0]
c = gl[gl[:, 0] == 2].astype(np.int64)
for nm, a, b in (("wait->inputs", 2, 4), ("inputs->pre-sync", 4, 5), ("cluster.sync", 5, 6), ("sync->stored", 6, 7), ("stored->exit", 7, 3), ("wait->exit", 2, 3)):
    d = (c[:, b] - c[:, a]) / 1e3
    print(f"cnorm {nm:18s} median {np.median(d):6.2f} even {np.median(d[0::2]):6.2f} odd {np.median(d[1::2]):6.2f}")
PY
