#!/bin/bash
mkdir -p gpurun_out
echo "### GEMM xmc correctness (timeout-wrapped)"
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_zz_gpu_new_formats.py -m gpu -q -p no:cacheprovider --timeout 120 -k "gemm" 2>&1 | tail -4
cat > /tmp/gk.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from blazr_b200 import ops, synth, decode
client = ops.B200Client(0)
for fmt, N, K, M in (("Q6_K", 14336, 4096, 2048), ("Q6_K", 28672, 4096, 2048), ("Q4_K", 14336, 4096, 4096), ("Q8_0", 14336, 4096, 2048)):
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
    print(f"XMC={os.environ.get('B200Q_GEMM_XMC','1')} NX={os.environ.get('B200Q_GEMM_NX')} NW={os.environ.get('B200Q_GEMM_NW')} {fmt} {N}x{K} M={M}: {us:.1f} us  {2.0*M*N*K/(us*1e-6)/1e12:.0f} TFLOP/s", flush=True)
    for w in ws: w.free()
PY
for cfg in "1 x x" "0 x x" "1 6 2" "1 5 3" "0 6 2"; do set -- $cfg; if [ "$2" = "x" ]; then B200Q_GEMM_XMC=$1 timeout 200 python /tmp/gk.py 2>&1 | grep XMC; else B200Q_GEMM_XMC=$1 B200Q_GEMM_NX=$2 B200Q_GEMM_NW=$3 timeout 200 python /tmp/gk.py 2>&1 | grep XMC; fi; done
echo "### launch list"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"matvec_kernel|attn_decode|add_rmsnorm|argmax|embed_kernel|swiglu" -c 900 --csv --log-file gpurun_out/r2_decode_step_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc $?"; wc -l gpurun_out/r2_decode_step_launches.csv
