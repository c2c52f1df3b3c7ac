"""torchrun script: expert-parallel MoE decode over N GPUs (DeepSeek-V2-Lite expert shapes) == the 1-GPU layer.

Placement (reference north_star: per-GPU expert placement; src/engine/executor_cache.rs:218-228 for the stacked layout):
the 64 routed experts + 2 shared halves are partitioned with shard_range, every rank builds ONLY its experts, runs its
local selected experts on the replicated hidden state (non-local slots masked to -1: the grouped kernel skips their weight
stream) and the partial outputs are summed by the one-shot NVLink all-reduce (b200q_allreduce: push to every peer, flags,
f64 sum in rank order -> identical bits on every rank).  Prints parity against rank 0's full single-GPU layer and the
per-layer time of the captured 5-launch sequence."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import decode, ops, synth, tp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
client = ops.B200Client(local)
dev = client.device
E, top_k, hidden, ffn, T = 66, 8, 2048, 1408, 1
tg, td = synth.GGML["Q4_K"], synth.GGML["Q8_0"]


def expert(e):
    gu = client.weight_from_ggml(tg, decode.random_ggml_device("Q4_K", 2 * ffn, hidden, 500 + e, dev), 2 * ffn, hidden)
    dn = client.weight_from_ggml(td, decode.random_ggml_device("Q8_0", hidden, ffn, 700 + e, dev), hidden, ffn)
    return ops.ExpertWeights(gu, dn, interleaved=True)


e0, e1 = tp.expert_range(E, rank, world)
moe = ops.MoeMlp(client, [expert(e) for e in range(e0, e1)], ffn, hidden, local=list(range(e0, e1)))
comm = ops.PeerComm(rank, world, T * hidden, dev)
g = torch.Generator(device="cpu"); g.manual_seed(3)
nsel = 16
sels = [torch.cat([torch.randperm(64, generator=g)[:6], torch.tensor([64, 65])]).to(torch.int32).reshape(T, top_k).to(dev) for _ in range(nsel)]
gw = torch.full((T, top_k), 1.0 / top_k, device=dev)
gx = torch.Generator(device=dev); gx.manual_seed(7)
x = torch.randn((T, hidden), device=dev, generator=gx)
outs = [torch.empty((T, hidden), device=dev) for _ in range(nsel)]


def layer(i):
    ls, lg = tp.ep_local_slots(sels[i], gw, e0, e1)
    part = moe.forward_decode(x, ls, lg)
    comm.allreduce(part, outs[i])


for i in range(nsel):
    layer(i)
torch.cuda.synchronize()
ok = True
if rank == 0:
    full = ops.MoeMlp(client, [expert(e) for e in range(E)], ffn, hidden)
    worst = 0.0
    for i in range(nsel):
        ref = full.forward_decode(x, sels[i], gw)
        err = float((outs[i] - ref).abs().max() / ref.abs().max())
        worst = max(worst, err)
    ok = worst < 1e-6
    print(f"ep{world}: {E} experts, {e1 - e0} on rank 0; max rel err vs the 1-GPU layer over {nsel} routings: {worst:.2e}", flush=True)
dist.barrier()
# every rank holds the same bits
chk = torch.stack([o.view(torch.int32).sum() for o in outs]).sum().reshape(1)
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(int(c) == int(allc[0]) for c in allc)
# timing: the 16 routings captured in one graph
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(gr, stream=s):
        for i in range(nsel):
            layer(i)
    for _ in range(3):
        gr.replay()
    s.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(10):
        gr.replay()
    b.record(s)
    s.synchronize()
us = torch.tensor([a.elapsed_time(b) * 1e3 / (10 * nsel)], device=dev)
dist.all_reduce(us, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"ep{world}: identical bits on every rank: {same}; MoE decode layer {float(us):.1f} us (max over ranks, graph replay, exchange included)", flush=True)
    print("EP_CHECK", "PASS" if ok and same else "FAIL", flush=True)
torch.cuda.synchronize(); sys.stdout.flush()
os._exit(0)
