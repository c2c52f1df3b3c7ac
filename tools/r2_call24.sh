#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -p no:cacheprovider --timeout 200 2>&1 | tail -4
cat > /tmp/gk.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from blazr_b200 import ops, synth, decode
client = ops.B200Client(0)
for fmt, N, K, M in (("Q6_K", 14336, 4096, 2048), ("Q6_K", 28672, 4096, 2048), ("Q4_K", 14336, 4096, 4096), ("Q8_0", 14336, 4096, 2048), ("Q6_K", 4096, 14336, 2048), ("Q6_K", 6144, 4096, 2048)):
    ws = [client.weight_from_ggml(synth.GGML[fmt], decode.random_ggml_device(fmt, N, K, 100 + i, client.device), N, K) for i in range(4)]
    x = torch.randn((M, K), device="cuda"); y = torch.empty((M, N), device="cuda")
    wss = [w.workspace(M) for w in ws]
    for w, s in zip(ws, wss): client.quant_matmul(x, w, out=y, workspace=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        for w, s in zip(ws, wss): client.quant_matmul(x, w, out=y, workspace=s)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"TAIL={os.environ.get('B200Q_GEMM_TAIL','1')} {fmt} {N}x{K} M={M}: {us:.1f} us  {2.0*M*N*K/(us*1e-6)/1e12:.0f} TFLOP/s", flush=True)
    for w in ws: w.free()
PY
for t in 1 0; do B200Q_GEMM_TAIL=$t timeout 300 python /tmp/gk.py 2>&1 | grep TAIL; done
