"""Kernel micro-benchmark: quant matvec GB/s per (format, shape, M) with CUDA-graph replay and L2 defeat.

Timing method: a CUDA graph holding one launch per rotating weight copy (copies
total > 2x L2 so every launch streams its weights from HBM), >= 20 warm-up replays, CUDA events on the
launching stream around R replays.  Algorithmic bytes = canonical packed weight bytes + x + y.

    python tools/kbench.py [--quick] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blazr_b200 import ops, synth  # noqa: E402

PEAK_GBS = 6544.0
try:
    PEAK_GBS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

L2_DEFEAT_BYTES = 300e6


def make_weight(client, fmt, N, K, seed):
    if fmt in synth.GGML:
        t = synth.GGML[fmt]
        return client.weight_from_ggml(t, synth.random_ggml(t, N, K, seed=seed), N, K)
    if fmt == "AWQ":
        qw, sc, zr = synth.random_awq(N, K, 128, seed=seed)
        return client.weight_from_decomposed(ops.DecomposedQuantTensor(qw, sc, zr, None, ops.DecomposedQuantMethod("awq", 128), (N, K)))
    if fmt == "GPTQ":
        qw, sc, qz, gi, _ = synth.random_gptq(N, K, 128, seed=seed)
        return client.weight_from_decomposed(ops.DecomposedQuantTensor(qw, sc, qz, gi, ops.DecomposedQuantMethod("gptq", 128), (N, K)))
    raise ValueError(fmt)


def bench_case(client, fmt, N, K, M, reps=20, path=ops.PATH_AUTO, split=True, chain=False):
    w0 = make_weight(client, fmt, N, K, seed=1)
    copies = int(min(96, max(2, -(-L2_DEFEAT_BYTES // w0.canonical_bytes))))
    ws = [w0] + [make_weight(client, fmt, N, K, seed=1) for _ in range(copies - 1)]  # same bytes, distinct buffers
    if chain:  # successor hints between the rotating copies (what a model's projection order gives the decode path)
        for a, b in zip(ws, ws[1:] + ws[:1]):
            a.set_next(b)
    x = torch.from_numpy(synth.random_act(M, K)).cuda()
    y = torch.empty((M, N), device="cuda", dtype=torch.float32)
    wss = [w.workspace(M) for w in ws]
    use_split = split and M <= 4 and path != ops.PATH_GEMM
    xq = client.quantize_act(x) if use_split else None

    def run_all():
        for w, s in zip(ws, wss):
            if use_split:
                client.matmul_q8(xq, M, w, out=y, workspace=s)
            else:
                client.quant_matmul(x, w, out=y, workspace=s, path=path)

    run_all()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            run_all()
        for _ in range(3):
            g.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps):
            g.replay()
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / (reps * copies)
    xb = x.element_size() if not use_split else 1.25
    bytes_alg = w0.canonical_bytes + M * K * xb + M * N * 4
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    flops = 2.0 * M * N * K
    res = dict(fmt=fmt, N=N, K=K, M=M, us=ms * 1e3, GBs=gbs, frac_hbm=gbs / PEAK_GBS, TFLOPs=flops / (ms * 1e-3) / 1e12, copies=copies,
               kernel_only=bool(use_split))
    for w in ws:
        w.free()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--fmts", default="Q4_K,Q6_K,Q8_0,AWQ")
    ap.add_argument("--ms", default="1")
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--chain", action="store_true", help="set b200q_weight_set_next between consecutive launches")
    args = ap.parse_args()
    client = ops.B200Client(0)
    shapes = [(4096, 4096), (14336, 4096), (4096, 14336), (28672, 4096)]
    if not args.quick:
        shapes = [(512, 2048), (2048, 2048), (8192, 2048), (1024, 4096)] + shapes + [(32000, 4096), (8192, 8192), (28672, 8192), (128256, 4096)]
    out = []
    for fmt in args.fmts.split(","):
        for (N, K) in shapes:
            for M in [int(m) for m in args.ms.split(",")]:
                r = bench_case(client, fmt, N, K, M, path=args.path, chain=args.chain)
                out.append(r)
                print(f"{fmt:5s} N={N:6d} K={K:6d} M={M:4d}  {r['us']:9.2f} us  {r['GBs']:8.1f} GB/s  {100 * r['frac_hbm']:5.1f}% HBM  "
                      f"{r['TFLOPs']:7.2f} TF/s  copies={r['copies']}", flush=True)
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
