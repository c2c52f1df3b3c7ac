"""ncu driver for the bench's roofline kernel: the gate|up matvec exactly as the decode step launches it (norm prologue +
SwiGLU epilogue on an interleaved gate|up weight), a few eager launches on rotating weight copies (no graphs, short)."""
import argparse, ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import decode, ops, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fmt", default="Q4_K"); ap.add_argument("--F", type=int, default=28672); ap.add_argument("--K", type=int, default=8192)
ap.add_argument("--iters", type=int, default=6); ap.add_argument("--plain", action="store_true", help="plain matmul_q8 instead of the fused variant")
a = ap.parse_args()
client = ops.B200Client(0)
dev = client.device
N = 2 * a.F
ws = [client.weight_from_ggml(synth.GGML[a.fmt], decode.random_ggml_device(a.fmt, N, a.K, 10 + i, dev), N, a.K) for i in range(3)]
L = ops.lib()
P = lambda t: C.c_void_p(t.data_ptr())
h = torch.randn((1, a.K), device=dev); h2 = torch.empty_like(h); wn = torch.ones(a.K, device=dev)
xq_ff = torch.zeros(int(L.b200q_act_bytes(C.c_int64(a.F), C.c_int64(1))), dtype=torch.uint8, device=dev)
xq = client.quantize_act(h)
y = torch.empty((1, N), device=dev)
for it in range(a.iters):
    w = ws[it % 3]
    s = w.workspace(1)
    if a.plain:
        client.matmul_q8(xq, 1, w, out=y, workspace=s)
    else:
        ops._check(L.b200q_matmul_norm_swiglu(w.handle, P(h), None, P(h2), P(wn), C.c_float(1e-5), C.c_int64(1), P(xq_ff), P(s), C.c_size_t(s.numel()), None))
torch.cuda.synchronize()
print("ok", int(xq_ff.sum()))
