"""Dual-format decode matvec (include/b200q.h b200q_weight_set_pair): a Q4_K weight and a Q6_K weight that read the same
activations -- q|k and v of a Q4_K_M file (reference src/loader/gguf.rs:365-372: every tensor keeps its own ggml type) -- are
computed by ONE stream-K launch.  Contract: the same bits as the two separate launches, for the plain and the norm-prologue
form, M = 1, 2, 4, ragged second weight, whole decode steps; unsupported pairings are refused with an error code."""
import ctypes as C

import numpy as np
import pytest
import torch

from blazr_b200 import decode, ops, synth

pytestmark = pytest.mark.gpu


def _weights(client, N1, N2, K, seed):
    t4, t6 = synth.GGML["Q4_K"], synth.GGML["Q6_K"]
    w1 = client.weight_from_ggml(t4, synth.random_ggml(t4, N1, K, seed=seed), N1, K)
    w2 = client.weight_from_ggml(t6, synth.random_ggml(t6, N2, K, seed=seed + 1), N2, K)
    return w1, w2


@pytest.mark.parametrize("N1,N2,K", [(256, 128, 1024), (1152, 128, 8192), (128 * 150, 1000, 512), (128 * 61, 128 * 37, 2048), (384, 200, 4096)])
@pytest.mark.parametrize("M", [1, 2, 4])
def test_dual_equals_two_launches(client, N1, N2, K, M):
    w1, w2 = _weights(client, N1, N2, K, seed=N1 + M)
    x = torch.from_numpy(synth.random_act(M, K, seed=5 + M)).cuda()
    xq = client.quantize_act(x)
    ref = torch.cat([client.matmul_q8(xq, M, w1), client.matmul_q8(xq, M, w2)], dim=1)
    w1.set_pair(w2)
    y = torch.zeros((M, N1 + N2), device="cuda")
    for _ in range(2):   # the workspace slots are their own flags: a second launch must find them clean
        y.zero_()
        client.matmul_q8(xq, M, w1, out=y)
        torch.cuda.synchronize()
        assert torch.equal(y, ref)
    # an output that cannot hold both weights is refused, never overrun
    with pytest.raises(ops.B200QError):
        client.matmul_q8(xq, M, w1, out=torch.zeros((M, N1), device="cuda"))
    w1.set_pair(None)
    assert torch.equal(client.matmul_q8(xq, M, w1), ref[:, :N1])


@pytest.mark.parametrize("N1,N2,K", [(1152, 128, 8192), (128 * 150, 640, 1024), (2560, 512, 2048)])
@pytest.mark.parametrize("M", [1, 4])
def test_dual_norm_prologue_equals_two_launches(client, N1, N2, K, M):
    w1, w2 = _weights(client, N1, N2, K, seed=N2 + M)
    g = torch.Generator(device="cuda"); g.manual_seed(K + M)
    h_in = torch.randn((M, K), device="cuda", generator=g)
    delta = torch.randn((M, K), device="cuda", generator=g)
    wn = 1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())

    def run(w, y, h_out):
        ws = w.workspace(M)
        ops._check(L.b200q_matmul_norm(w.handle, P(h_in), P(delta), P(h_out) if h_out is not None else None, P(wn), C.c_float(1e-5), C.c_int64(M), P(y),
                                       C.c_int32(ops.F32), C.c_int64(y.stride(0)), P(ws), C.c_size_t(ws.numel()), None))

    ref = torch.zeros((M, N1 + N2), device="cuda"); h_ref = torch.zeros_like(h_in)
    run(w1, ref, h_ref)
    run(w2, ref[:, N1:], None)
    torch.cuda.synchronize()
    w1.set_pair(w2)
    y = torch.zeros((M, N1 + N2), device="cuda"); h_out = torch.zeros_like(h_in)
    run(w1, y, h_out)
    torch.cuda.synchronize()
    assert torch.equal(h_out, h_ref) and torch.equal(y, ref)


def test_unsupported_pairings_are_refused(client):
    t4, t6, t8 = synth.GGML["Q4_K"], synth.GGML["Q6_K"], synth.GGML["Q8_0"]
    mk = lambda t, N, K, s: client.weight_from_ggml(t, synth.random_ggml(t, N, K, seed=s), N, K)
    a, b = mk(t4, 256, 1024, 1), mk(t6, 128, 1024, 2)
    with pytest.raises(ops.B200QError):
        mk(t4, 200, 1024, 3).set_pair(b)            # first weight not tile-aligned
    with pytest.raises(ops.B200QError):
        a.set_pair(mk(t6, 128, 2048, 4))            # different K
    with pytest.raises(ops.B200QError):
        a.set_pair(mk(t8, 128, 1024, 5))            # no Q4_K + Q8_0 kernel: launch them separately
    with pytest.raises(ops.B200QError):
        b.set_pair(a)                               # order matters (Q6_K first is not built)
    a.set_pair(b)
    with pytest.raises(ops.B200QError):
        mk(t4, 256, 1024, 6).set_pair(a)            # a weight that has a partner cannot be one


@pytest.mark.parametrize("scheme", ["Q4_K_M"])
def test_decode_with_dual_launch_reproduces_separate_launches(client, scheme, monkeypatch):
    """whole decode steps (graph replay): the dual q|k + v launch changes no bit of the logits nor of the greedy stream"""
    cfg = decode.ModelConfig("dual-8", 1024, 8, 8, 2, 128, 2048, 4096, 10000.0)
    hm = decode.build_host_model(cfg, scheme, seed=11)
    prompt = np.asarray([[5, 900, 33, 2047, 1]])
    monkeypatch.setenv("B200Q_DUAL", "0")
    a = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    monkeypatch.setenv("B200Q_DUAL", "1")
    b = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    assert any(len(l["qkv_mv"]) < len(l["qkv"]) for l in b.layers), "the preset has no Q4_K + Q6_K q|k|v layer"
    assert all(len(l["qkv_mv"]) == len(l["qkv"]) for l in a.layers)
    ta, tb = a.generate(prompt, 24, use_graph=True), b.generate(prompt, 24, use_graph=True)
    assert np.array_equal(ta, tb)
    assert np.array_equal(a.full_logits().cpu().numpy().view(np.uint32), b.full_logits().cpu().numpy().view(np.uint32))
    assert b.launches_per_step() < a.launches_per_step()
