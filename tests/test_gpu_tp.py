"""Fused tensor-parallel exchange (csrc/comm_dev.cuh) on hardware.

World 1 (the driver's single-GPU box): the row-parallel matvec pushes its f64 row sums into the exchange slots instead of
y, the consumers (stand-alone finish, fused add+RMSNorm+quantise, gathered arg-max) read them back -- every result must
equal, bit for bit, the plain matmul_q8 / add_rmsnorm_quant / argmax path, over repeated calls (epoch / parity protocol)
and under CUDA-graph replay.

World > 1 (only where the box has >= 2 GPUs): tools/tp_check.py under torchrun -- tensor-parallel decode (column-split
q/k/v/gate/up, row-split o/down with the fused exchange, vocabulary-split lm_head) reproduces the 1-GPU logits bit for
bit and the same greedy stream.  Reference: src/engine/tensor_parallel.rs:61-163.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from blazr_b200 import ops, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fmt,N,K", [("Q4_K", 1024, 2048), ("Q6_K", 640, 1024), ("Q8_0", 512, 4096)])
@pytest.mark.parametrize("M", [1, 2, 4])
def test_rowpar_push_then_finish_equals_plain_matvec(client, fmt, N, K, M):
    t = synth.GGML[fmt]
    w = client.weight_from_ggml(t, synth.random_ggml(t, N, K, seed=5), N, K)
    x = torch.from_numpy(synth.random_act(M, K, seed=6)).cuda()
    xq = client.quantize_act(x)
    y32 = client.matmul_q8(xq, M, w)
    comm = ops.PeerComm(0, 1, M * N, client.device)
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    for _ in range(5):  # alternating parities
        out.zero_()
        comm.matmul_q8_rowpar(w, xq, M, N, w.workspace(M))
        comm.allreduce_finish(out)
        assert torch.equal(out, y32)
    # mixed with the stand-alone all-reduce on the same buffers (same epochs / flags)
    comm.allreduce(y32, out)
    assert torch.equal(out, y32)
    comm.matmul_q8_rowpar(w, xq, M, N, w.workspace(M))
    comm.allreduce_finish(out)
    assert torch.equal(out, y32)
    comm.free()


@pytest.mark.parametrize("H", [512, 2048, 4096, 8192])
@pytest.mark.parametrize("M", [1, 3])
def test_add_rmsnorm_quant_equals_oracle(client, H, M):
    """add+RMSNorm+quantise against the CPU restatement, bit for bit (the fused-exchange consumer below is then checked
    against this kernel)"""
    import ctypes as C
    g = torch.Generator(device="cuda"); g.manual_seed(H + M)
    h_in = torch.randn((M, H), device="cuda", generator=g)
    delta = torch.randn((M, H), device="cuda", generator=g)
    wn = 1.0 + 0.1 * torch.randn(H, device="cuda", generator=g)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())
    nb = int(L.b200q_act_bytes(C.c_int64(H), C.c_int64(M)))
    h_out = torch.empty_like(h_in); xq = torch.zeros(nb, dtype=torch.uint8, device="cuda"); xn = torch.empty_like(h_in)
    ops._check(L.b200q_add_rmsnorm_quant(P(h_in), P(delta), P(h_out), P(wn), C.c_float(1e-5), C.c_int64(H), C.c_int64(M), P(xq), P(xn), None))
    torch.cuda.synchronize()
    # the CPU oracle's restatement (oracle/model.py rmsnorm: f32 add, f64 sum of squares, f32 1/sqrt, two f32 multiplies)
    import oracle
    from oracle.model import rmsnorm
    h = (h_in.cpu().numpy() + delta.cpu().numpy()).astype(np.float32)
    assert np.array_equal(h_out.cpu().numpy().view(np.uint32), h.view(np.uint32))
    xref = np.stack([rmsnorm(h[m], wn.cpu().numpy(), 1e-5) for m in range(M)])
    assert np.array_equal(xn.cpu().numpy().view(np.uint32), xref.view(np.uint32))
    q, d, _ = client.act_unpack(xq, M, H)
    qr, dr, _ = oracle.quantize_act(xref)
    assert np.array_equal(q.cpu().numpy(), qr) and np.array_equal(d.cpu().numpy().view(np.uint32), dr.view(np.uint32))


@pytest.mark.parametrize("H", [512, 2048, 4096, 8192])
def test_fused_exchange_norm_consumer_world1(client, H):
    """producer (row-parallel matvec -> slots) + consumer (allreduce + add + RMSNorm + quantise) == matvec, then
    add_rmsnorm_quant with that delta; eager and under graph replay"""
    import ctypes as C
    M, K = 2, 1024
    t = synth.GGML["Q4_K"]
    w = client.weight_from_ggml(t, synth.random_ggml(t, H, K, seed=9), H, K)
    x = torch.from_numpy(synth.random_act(M, K, seed=10)).cuda()
    xq_in = client.quantize_act(x)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    h_in = torch.randn((M, H), device="cuda", generator=g)
    wn = 1.0 + 0.1 * torch.randn(H, device="cuda", generator=g)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())
    nb = int(L.b200q_act_bytes(C.c_int64(H), C.c_int64(M)))
    delta = client.matmul_q8(xq_in, M, w)
    h_ref = torch.empty_like(h_in); xq_ref = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ops._check(L.b200q_add_rmsnorm_quant(P(h_in), P(delta), P(h_ref), P(wn), C.c_float(1e-5), C.c_int64(H), C.c_int64(M), P(xq_ref), None, None))
    comm = ops.PeerComm(0, 1, M * H, client.device)
    h_out = torch.empty_like(h_in); xq = torch.zeros(nb, dtype=torch.uint8, device="cuda")

    def run():
        comm.matmul_q8_rowpar(w, xq_in, M, H, w.workspace(M))
        comm.allreduce_add_rmsnorm_quant(h_in, h_out, wn, 1e-5, H, M, xq=xq)

    for _ in range(3):
        h_out.zero_(); xq.zero_()
        run()
        assert torch.equal(h_out, h_ref) and torch.equal(xq, xq_ref)
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            run()
    torch.cuda.synchronize()
    for _ in range(4):
        h_out.zero_(); xq.zero_()
        gr.replay()
        torch.cuda.synchronize()
        assert torch.equal(h_out, h_ref) and torch.equal(xq, xq_ref)
    comm.free()


def test_gather_push_then_argmax_world1(client):
    M, N, K = 2, 1000, 1024   # N not a multiple of 128: padding columns of the region stay -inf
    vs = 1024
    t = synth.GGML["Q6_K"]
    w = client.weight_from_ggml(t, synth.random_ggml(t, N, K, seed=3), N, K)
    xq = client.quantize_act(torch.from_numpy(synth.random_act(M, K, seed=4)).cuda())
    y = client.matmul_q8(xq, M, w)
    comm = ops.PeerComm(0, 1, 16, client.device, gather_elems=M * vs)
    ids = torch.zeros(M, dtype=torch.int64, device="cuda")
    pos = torch.zeros(M, dtype=torch.int32, device="cuda")
    for it in range(3):
        comm.matmul_q8_gather(w, xq, M, vs, w.workspace(M))
        comm.argmax_gathered(vs, M, ids, pos)
        torch.cuda.synchronize()
        got = comm.gathered()[0, :M * vs].reshape(M, vs)
        assert torch.equal(got[:, :N], y) and bool(torch.isinf(got[:, N:]).all())
        assert torch.equal(ids, y.argmax(dim=1)) and int(pos[0]) == it + 1
    comm.free()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tensor_parallel_decode_matches_one_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs on one box (gpurun --gpus {world})")
    port = 29600 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tools", "tp_check.py")], capture_output=True, text=True, timeout=600)
    tail = "\n".join(l for l in (r.stdout + r.stderr).splitlines() if l.startswith("tp") or "TP_CHECK" in l or "Error" in l)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"tp_check_world{world}.log"), "w").write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert "TP_CHECK PASS" in r.stdout, tail
