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
emu = int(os.environ.get("B200Q_EMULATE_TP", "0"))
dec = (decode.Decoder(client, decode.PRESETS[model], scheme, batch=1, max_ctx=256, tp_rank=0, tp_world=emu, emulate_shard=True) if emu > 1
       else decode.Decoder(client, decode.PRESETS[model], scheme, batch=1, max_ctx=256))
dec.reset([1]); dec.pos.fill_(48)
dec.step(); torch.cuda.synchronize()
n_launch = sum(len(l["qkv_mv"]) + len(l["o"]) + len(l["gu"]) + len(l["down"]) for l in dec.layers) + len(dec.head)
# capture() first runs the step eagerly (launch indices 0 .. n_launch-1), then captures it (n_launch .. 2 n_launch - 1): the
# graph replays stamp the SECOND half (round 2's first version read the eager half: CPU launch gaps, not the graph's timeline)
trace = torch.zeros(2 * n_launch * 148 * 8, dtype=torch.int64, device="cuda")
ops.lib().b200q_debug_set_matvec_trace(C.c_void_p(trace.data_ptr()))
dec.graph = None
dec.capture()
ops.lib().b200q_debug_set_matvec_trace(None)
for _ in range(3):
    dec.graph.replay()
torch.cuda.synchronize()
# glue kernels (attention, add+norm; TRACE build): stamped during one more replay, merged into the timeline below
GCAP = 4096
glue = torch.zeros(GCAP * 8, dtype=torch.int64, device="cuda")
rc_glue = ops.lib().b200q_debug_set_glue_trace(C.c_void_p(glue.data_ptr()), C.c_int32(GCAP)) if hasattr(ops.lib(), "b200q_debug_set_glue_trace") else -99
has_glue = rc_glue == 0
dec.graph.replay()
torch.cuda.synchronize()
if has_glue:
    ops.lib().b200q_debug_set_glue_trace(None, C.c_int32(0))
gl = glue.cpu().numpy().reshape(GCAP, 8).astype(np.float64)
gl = gl[gl[:, 1] > 0]
print(f"# glue trace: rc {rc_glue}, {len(gl)} records")
t = trace.cpu().numpy().reshape(2 * n_launch, 148, 8)[n_launch:].astype(np.float64)
if os.environ.get("B200Q_TRACE_DUMP"):   # raw stamps for offline analysis (int64 ns)
    np.savez_compressed(os.environ["B200Q_TRACE_DUMP"], matvec=trace.cpu().numpy().reshape(2 * n_launch, 148, 8)[n_launch:], glue=glue.cpu().numpy().reshape(GCAP, 8))
# launch order of the matvec kernels inside one step (mixed-format groups are several launches)
names = []
for li, l in enumerate(dec.layers):
    for key in ("qkv_mv", "o", "gu", "down"):
        for j, _ in enumerate(l[key]):
            names.append((li, key.replace("_mv", "") if len(l[key]) == 1 else f"{key.replace('_mv', '')}{j}"))
names += [(-1, "head")] * len(dec.head)
t0 = None
prev_exit = None
prev_wait_abs = None
first = min(2, len(dec.layers) - a.layers)
for k, (li, nm) in enumerate(names):
    tk = t[k]; tk = tk[tk[:, 0] > 0]
    if len(tk) == 0:
        continue
    if li < first or li >= first + a.layers:
        prev_wait_abs = np.median(tk[:, 4])
        continue
    if t0 is None: t0 = tk[:, 0].min()
    rel = (tk - t0) / 1e3
    ent, wait, data, ex = rel[:, 0], rel[:, 4], rel[:, 1], rel[:, 3]
    fx = rel[:, 7][rel[:, 7] > -1e6]
    done = max(ex.max(), fx.max() if len(fx) else 0)
    gap = f" since_prev_done {np.median(wait) - prev_exit:6.2f}" if prev_exit is not None else ""
    if t0 is None: t0 = tk[:, 0].min()
    if has_glue:   # glue kernels that started between the previous matvec's first entry and this one's wait
        lo = -1e30 if prev_wait_abs is None else prev_wait_abs
        for g_ in gl[(gl[:, 1] > lo) & (gl[:, 1] <= np.median(tk[:, 4]))]:
            kind = {1: "attn", 2: "cnorm", 3: "norm"}.get(int(g_[0]), "?")
            print(f"      {kind:5s}          entry {(g_[1] - t0) / 1e3:7.2f}                   wait {(g_[2] - t0) / 1e3:7.2f}              exit {(g_[3] - t0) / 1e3:7.2f}")
    prev_wait_abs = np.median(tk[:, 4])
    print(f"L{li:02d} {nm:5s} ctas {len(tk):3d} entry {np.median(ent):7.2f} [{ent.min():7.2f},{ent.max():7.2f}] wait {np.median(wait):7.2f} data {np.median(data):7.2f} exit {np.median(ex):7.2f} max {ex.max():7.2f} done {done:7.2f} | busy {done - np.median(wait):6.2f}{gap}")
    prev_exit = done
