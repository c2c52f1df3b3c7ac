"""Debug: globaltimer trace of every matvec launch inside one captured decode step."""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import decode, ops
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="mistral-7b:Q6_K"); ap.add_argument("--layers", type=int, default=3)
a = ap.parse_args()
model, scheme = a.workload.split(":")
client = ops.B200Client(0)
dec = decode.Decoder(client, decode.PRESETS[model], scheme, batch=1, max_ctx=256)
dec.reset([1]); dec.pos.fill_(48)
dec.step(); torch.cuda.synchronize()
n_launch = sum(len(l["qkv"]) + len(l["o"]) + len(l["gu"]) + len(l["down"]) for l in dec.layers) + len(dec.head)
trace = torch.zeros(n_launch * 148 * 8, dtype=torch.int64, device="cuda")
ops.lib().b200q_debug_set_matvec_trace(C.c_void_p(trace.data_ptr()))
dec.graph = None
dec.capture()
ops.lib().b200q_debug_set_matvec_trace(None)
for _ in range(3):
    dec.graph.replay()
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(n_launch, 148, 8).astype(np.float64)
names = ["qkv", "o", "gu", "down"]
t0 = None
prev_exit = None
for k in range(4 * 2, 4 * (2 + a.layers)):
    tk = t[k]; tk = tk[tk[:, 0] > 0]
    if t0 is None: t0 = tk[:, 0].min()
    rel = (tk - t0) / 1e3
    ent, wait, data, ex = rel[:, 0], rel[:, 4], rel[:, 1], rel[:, 3]
    fx = rel[:, 7][rel[:, 7] > -1e6]
    done = max(ex.max(), fx.max() if len(fx) else 0)
    gap = f" gap_since_prev_done {ent.min() - prev_exit:6.2f}" if prev_exit is not None else ""
    print(f"{names[k % 4]:5s} entry {np.median(ent):7.2f} [{ent.min():7.2f},{ent.max():7.2f}] wait {np.median(wait):7.2f} data {np.median(data):7.2f} exit {np.median(ex):7.2f} max {ex.max():7.2f} done {done:7.2f} | busy {done - np.median(wait):6.2f}{gap}")
    prev_exit = done

# detail of one gate|up launch: slowest CTAs
k = 4 * 3 + 2
tk = t[k]
rel = (tk - t0) / 1e3
order = np.argsort(rel[:, 3])
print("gu launch", k, ": CTAs sorted by exit (g, smid, entry, data, exit); first 5 and last 12")
for g_ in list(order[:5]) + list(order[-12:]):
    print(f"  g={g_:3d} sm={int(tk[g_,2]):3d} entry {rel[g_,0]:7.2f} data {rel[g_,1]:7.2f} exit {rel[g_,3]:7.2f}")
# which other kernels' CTAs were on the same SMs? compare with the next launch (down)
tn = t[k + 1]; reln = (tn - t0) / 1e3
sm_next = {int(tn[g_, 2]): reln[g_, 0] for g_ in range(148)}
late = [int(tk[g_, 2]) for g_ in order[-12:]]
print("entry time of the *down* CTA on those SMs:", [round(sm_next.get(sm, -1), 1) for sm in late])
print("down entries (sorted):", np.round(np.sort(reln[:, 0])[[0, 37, 74, 111, 147]], 1))
