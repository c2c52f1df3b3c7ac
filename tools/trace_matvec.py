"""Debug: per-CTA globaltimer trace of a PDL chain of matvec kernels replayed from a CUDA graph."""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import ops, synth
from tools.kbench import make_weight

ap = argparse.ArgumentParser()
ap.add_argument("--fmt", default="Q6_K"); ap.add_argument("--N", type=int, default=14336); ap.add_argument("--K", type=int, default=4096)
ap.add_argument("--n", type=int, default=8); ap.add_argument("--chain", action="store_true")
a = ap.parse_args()
client = ops.B200Client(0)
ws = [make_weight(client, a.fmt, a.N, a.K, seed=1) for _ in range(a.n)]
if a.chain:
    for w0_, w1_ in zip(ws, ws[1:] + ws[:1]):
        w0_.set_next(w1_)
x = torch.from_numpy(synth.random_act(1, a.K)).cuda()
xq = client.quantize_act(x)
y = torch.empty((1, a.N), device="cuda")
wss = [w.workspace(1) for w in ws]
trace = torch.zeros(a.n * 148 * 8, dtype=torch.int64, device="cuda")
for w, s in zip(ws, wss):
    client.matmul_q8(xq, 1, w, out=y, workspace=s)
torch.cuda.synchronize()
ops.lib().b200q_debug_set_matvec_trace(C.c_void_p(trace.data_ptr()))
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    with torch.cuda.graph(g, stream=st):
        for w, s in zip(ws, wss):
            client.matmul_q8(xq, 1, w, out=y, workspace=s)
    ops.lib().b200q_debug_set_matvec_trace(None)
    for _ in range(3):
        g.replay()
    st.synchronize()
t = trace.cpu().numpy().reshape(a.n, 148, 8).astype(np.float64)
t0 = t[0, :, 0][t[0, :, 0] > 0].min()
print(f"{a.fmt} {a.N}x{a.K}: per-kernel timeline (us since first CTA entry of kernel 0); med [min,max] over CTAs")
prev_exit = None
for k in range(a.n):
    tk = t[k]
    tk = tk[tk[:, 0] > 0]
    rel = (tk - t0) / 1e3
    line = f"k{k}: "
    for nm, col in (("entry", 0), ("wait", 4), ("data", 1), ("exit", 3), ("fxbar", 5), ("fxatom", 6), ("fxdone", 7)):
        v = rel[:, col]
        v = v[v > -1e6]
        if len(v) == 0:
            continue
        line += f"{nm} {np.median(v):6.2f} [{v.min():6.2f},{v.max():6.2f}]  "
    ex = rel[:, 3].max()
    if prev_exit is not None:
        line += f" period {ex - prev_exit:5.2f}"
    prev_exit = ex
    print(line)
