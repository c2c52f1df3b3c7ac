"""Tiny driver for ncu: a few eager launches of one matmul shape (no graphs, short)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import ops, synth  # noqa: E402
from tools.kbench import make_weight  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fmt", default="Q6_K")
ap.add_argument("--N", type=int, default=14336)
ap.add_argument("--K", type=int, default=4096)
ap.add_argument("--M", type=int, default=1)
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--path", type=int, default=0)
a = ap.parse_args()
client = ops.B200Client(0)
ws = [make_weight(client, a.fmt, a.N, a.K, seed=1) for _ in range(4)]
x = torch.from_numpy(synth.random_act(a.M, a.K)).cuda()
for it in range(a.iters):
    y = client.quant_matmul(x, ws[it % 4], path=a.path)
torch.cuda.synchronize()
print("ok", float(y.float().abs().sum()))
