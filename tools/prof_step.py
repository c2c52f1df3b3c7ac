"""ncu driver: a few eager decode steps of a model preset (kernels only; no graph)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import decode, ops
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="mistral-7b:Q6_K"); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--ctx", type=int, default=48)
a = ap.parse_args()
model, scheme = a.workload.split(":")
client = ops.B200Client(0)
dec = decode.Decoder(client, decode.PRESETS[model], scheme, batch=1, max_ctx=256)
dec.reset([1])
dec.pos.fill_(a.ctx)
for _ in range(a.steps):
    dec.step()
torch.cuda.synchronize()
print("ok", int(dec.ids[0]), dec.launches_per_step())
