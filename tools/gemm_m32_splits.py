import os, sys, torch
sys.path.insert(0, os.getcwd())
from blazr_b200 import ops, synth, decode
client = ops.B200Client(0)
M = 32
for fmt, N, K in (("Q6_K", 14336, 4096), ("Q6_K", 4096, 14336), ("Q6_K", 6144, 4096), ("Q6_K", 4096, 4096), ("Q4_K", 28672, 4096)):
    ws = [client.weight_from_ggml(synth.GGML[fmt], decode.random_ggml_device(fmt, N, K, 100 + i, client.device), N, K) for i in range(6)]
    x = torch.randn((M, K), device="cuda"); y = torch.empty((M, N), device="cuda")
    wss = [w.workspace(M) for w in ws]
    for w, s in zip(ws, wss): client.quant_matmul(x, w, out=y, workspace=s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for w, s in zip(ws, wss): client.quant_matmul(x, w, out=y, workspace=s)
        for _ in range(3): g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10): g.replay()
        e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 60
    print(f"SPLITS={os.environ.get('B200Q_GEMM_SPLITS')} {fmt} {N}x{K} M={M}: {us:.1f} us  {ws[0].canonical_bytes/(us*1e-6)/1e9:.0f} GB/s", flush=True)
    for w in ws: w.free()
