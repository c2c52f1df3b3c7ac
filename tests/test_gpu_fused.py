"""Fused activation producers of the decode matvec (include/b200q.h b200q_matmul_norm / b200q_matmul_swiglu): the consumer
warps build the quantised activation in shared memory, so no separate norm / SwiGLU launch runs.  Contract: the same bits
as the separate operators followed by b200q_matmul_q8 (identical arithmetic), at kernel level and through whole decode steps."""
import ctypes as C

import numpy as np
import pytest
import torch

from blazr_b200 import decode, ops, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", ["Q4_K", "Q6_K", "Q8_0"])
@pytest.mark.parametrize("K", [512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("M", [1, 2, 4])
def test_matmul_norm_equals_norm_then_matvec(client, fmt, K, M):
    N = 384
    t = synth.GGML[fmt]
    w = client.weight_from_ggml(t, synth.random_ggml(t, N, K, seed=K + M), N, K)
    g = torch.Generator(device="cuda"); g.manual_seed(K * 7 + M)
    h_in = torch.randn((M, K), device="cuda", generator=g)
    delta = torch.randn((M, K), device="cuda", generator=g)
    wn = 1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())
    nb = int(L.b200q_act_bytes(C.c_int64(K), C.c_int64(M)))
    for dl in (delta, None):
        h_ref = torch.empty_like(h_in); xq = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        ops._check(L.b200q_add_rmsnorm_quant(P(h_in), P(dl) if dl is not None else None, P(h_ref), P(wn), C.c_float(1e-5), C.c_int64(K), C.c_int64(M), P(xq), None, None))
        y_ref = client.matmul_q8(xq, M, w)
        h_out = torch.zeros_like(h_in); y = torch.zeros((M, N), device="cuda")
        ws = w.workspace(M)
        ops._check(L.b200q_matmul_norm(w.handle, P(h_in), P(dl) if dl is not None else None, P(h_out), P(wn), C.c_float(1e-5), C.c_int64(M), P(y),
                                       C.c_int32(ops.F32), C.c_int64(N), P(ws), C.c_size_t(ws.numel()), None))
        torch.cuda.synchronize()
        assert torch.equal(h_out, h_ref)
        assert torch.equal(y, y_ref)


@pytest.mark.parametrize("fmt,K,N,M", [("Q4_K", 512, 128 * 350, 1), ("Q4_K", 1024, 128 * 201, 2), ("Q6_K", 2048, 128 * 130, 1), ("Q8_0", 1024, 128 * 263, 4)])
def test_matmul_norm_stream_k_head_tail_and_full_segments(client, fmt, K, N, M):
    """more chunks than CTAs with a ragged split: a CTA's range holds a head segment (k-chunks kcH..KC-1 of its first tile), a
    tail segment (k-chunks 0..nT-1 of its last tile) AND full tiles, so the in-shared-memory activation records are re-walked
    from k-chunk 0 at every segment switch (regression: the record offset was not rewound after a tail segment)"""
    t = synth.GGML[fmt]
    w = client.weight_from_ggml(t, synth.random_ggml(t, N, K, seed=K + M), N, K)
    g = torch.Generator(device="cuda"); g.manual_seed(K * 3 + M)
    h_in = torch.randn((M, K), device="cuda", generator=g)
    delta = torch.randn((M, K), device="cuda", generator=g)
    wn = 1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())
    nb = int(L.b200q_act_bytes(C.c_int64(K), C.c_int64(M)))
    h_ref = torch.empty_like(h_in); xq = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ops._check(L.b200q_add_rmsnorm_quant(P(h_in), P(delta), P(h_ref), P(wn), C.c_float(1e-5), C.c_int64(K), C.c_int64(M), P(xq), None, None))
    y_ref = client.matmul_q8(xq, M, w)
    h_out = torch.zeros_like(h_in); y = torch.zeros((M, N), device="cuda")
    ws = w.workspace(M)
    ops._check(L.b200q_matmul_norm(w.handle, P(h_in), P(delta), P(h_out), P(wn), C.c_float(1e-5), C.c_int64(M), P(y),
                                   C.c_int32(ops.F32), C.c_int64(N), P(ws), C.c_size_t(ws.numel()), None))
    torch.cuda.synchronize()
    assert torch.equal(h_out, h_ref)
    assert torch.equal(y, y_ref)


def test_matmul_norm_rejects_unsupported_k(client):
    t = synth.GGML["Q8_0"]
    K = 768  # not 512 * 2^j
    w = client.weight_from_ggml(t, synth.random_ggml(t, 128, K, seed=1), 128, K)
    h = torch.zeros((1, K), device="cuda"); y = torch.zeros((1, 128), device="cuda"); ws = w.workspace(1)
    P = lambda t_: C.c_void_p(t_.data_ptr())
    rc = ops.lib().b200q_matmul_norm(w.handle, P(h), None, None, P(h), C.c_float(1e-5), C.c_int64(1), P(y), C.c_int32(ops.F32), C.c_int64(128), P(ws),
                                     C.c_size_t(ws.numel()), None)
    assert rc == -2  # B200Q_ERR_UNSUPPORTED: an error code, never a silent fallback


@pytest.mark.parametrize("scheme", ["Q4_K_M", "Q6_K", "AWQ"])
def test_decode_with_fused_norm_reproduces_the_separate_operators(client, scheme, monkeypatch):
    cfg = decode.PRESETS["tiny"]
    hm = decode.build_host_model(cfg, scheme, seed=17)
    monkeypatch.setenv("B200Q_FUSED", "0")
    ref = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    monkeypatch.setenv("B200Q_FUSED", "1")
    fus = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    assert fus.fused and not ref.fused
    for d in (ref, fus):
        d.reset([7])
    for _ in range(6):
        ref.step(); fus.step()
        torch.cuda.synchronize()
        assert np.array_equal(ref.logits.cpu().numpy().view(np.uint32), fus.logits.cpu().numpy().view(np.uint32))
    prompt = np.asarray([[3, 1, 4, 1, 5]])
    assert np.array_equal(ref.generate(prompt, 16, use_graph=True), fus.generate(prompt, 16, use_graph=True))
    assert fus.launches_per_step() < ref.launches_per_step()


@pytest.mark.parametrize("fmt", ["Q4_K", "Q6_K", "Q8_0", "Q5_K"])
@pytest.mark.parametrize("F,K", [(64, 512), (704, 1024), (1408, 2048), (4096, 4096)])
@pytest.mark.parametrize("M", [1, 2, 4])
def test_swiglu_epilogue_equals_matvec_then_swiglu(client, fmt, F, K, M):
    """gate|up uploaded in the interleaved row order + fused epilogue == plain gate|up matvec + swiglu_quant, record bytes equal;
    F = 704 / 1408 leave the last 256-k record chunk partly unwritten (zero padding), many tiles are split between CTAs"""
    t = synth.GGML[fmt]
    blocks = synth.random_ggml(t, 2 * F, K, seed=F + K + M)          # rows [gate; up]
    w_plain = client.weight_from_ggml(t, blocks, 2 * F, K)
    w_il = client.weight_from_ggml(t, np.ascontiguousarray(blocks[ops.gate_up_row_order(F)]), 2 * F, K)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())
    perm = ops.gate_up_row_order(F)
    assert all(int(L.b200q_gate_up_row(C.c_int64(F), C.c_int64(r))) == int(perm[r]) for r in (0, 3, 4, 7, 8, 127, 2 * F - 1))
    x = torch.from_numpy(synth.random_act(M, K, seed=11)).cuda()
    xq = client.quantize_act(x)
    gu = client.matmul_q8(xq, M, w_plain)
    nb = int(L.b200q_act_bytes(C.c_int64(F), C.c_int64(M)))
    ref = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ops._check(L.b200q_swiglu_quant(P(gu), C.c_int64(F), C.c_int64(M), P(ref), None))
    out = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ws = w_il.workspace(M)
    for _ in range(2):
        out.zero_()
        ops._check(L.b200q_matmul_q8_swiglu(w_il.handle, P(xq), C.c_int64(M), P(out), P(ws), C.c_size_t(ws.numel()), None))
        torch.cuda.synchronize()
        assert torch.equal(out, ref)


def test_norm_swiglu_both_fused(client):
    """norm prologue + SwiGLU epilogue in one launch == the four separate operators"""
    t = synth.GGML["Q4_K"]
    F, K, M = 1024, 2048, 2
    blocks = synth.random_ggml(t, 2 * F, K, seed=3)
    w_plain = client.weight_from_ggml(t, blocks, 2 * F, K)
    w_il = client.weight_from_ggml(t, np.ascontiguousarray(blocks[ops.gate_up_row_order(F)]), 2 * F, K)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    h_in = torch.randn((M, K), device="cuda", generator=g); delta = torch.randn((M, K), device="cuda", generator=g)
    wn = 1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)
    L = ops.lib()
    P = lambda t_: C.c_void_p(t_.data_ptr())
    xq = torch.zeros(int(L.b200q_act_bytes(C.c_int64(K), C.c_int64(M))), dtype=torch.uint8, device="cuda")
    h_ref = torch.empty_like(h_in)
    ops._check(L.b200q_add_rmsnorm_quant(P(h_in), P(delta), P(h_ref), P(wn), C.c_float(1e-5), C.c_int64(K), C.c_int64(M), P(xq), None, None))
    gu = client.matmul_q8(xq, M, w_plain)
    nb = int(L.b200q_act_bytes(C.c_int64(F), C.c_int64(M)))
    ref = torch.zeros(nb, dtype=torch.uint8, device="cuda")
    ops._check(L.b200q_swiglu_quant(P(gu), C.c_int64(F), C.c_int64(M), P(ref), None))
    out = torch.zeros(nb, dtype=torch.uint8, device="cuda"); h_out = torch.zeros_like(h_in)
    ws = w_il.workspace(M)
    ops._check(L.b200q_matmul_norm_swiglu(w_il.handle, P(h_in), P(delta), P(h_out), P(wn), C.c_float(1e-5), C.c_int64(M), P(out), P(ws), C.c_size_t(ws.numel()), None))
    torch.cuda.synchronize()
    assert torch.equal(h_out, h_ref) and torch.equal(out, ref)


@pytest.mark.parametrize("scheme", ["Q4_K_M", "Q6_K", "Q8_0"])
def test_decode_default_fusions_reproduce_the_separate_operators(client, scheme, monkeypatch):
    """the default decode step (norm prologue + SwiGLU epilogue: 5 launches per layer) gives the logits of the 8-launch step, bit for bit"""
    cfg = decode.PRESETS["tiny"]
    hm = decode.build_host_model(cfg, scheme, seed=23)
    monkeypatch.setenv("B200Q_FUSED", "0"); monkeypatch.setenv("B200Q_SWIGLU_EPI", "0")
    ref = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    monkeypatch.delenv("B200Q_FUSED"); monkeypatch.delenv("B200Q_SWIGLU_EPI")
    fus = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    assert fus.fused and fus.layers[0]["swiglu_epi"] and not ref.fused and not ref.layers[0]["swiglu_epi"]
    for d in (ref, fus):
        d.reset([7])
    for _ in range(6):
        ref.step(); fus.step()
        torch.cuda.synchronize()
        assert np.array_equal(ref.logits.cpu().numpy().view(np.uint32), fus.logits.cpu().numpy().view(np.uint32))
    prompt = np.asarray([[3, 1, 4, 1, 5]])
    assert np.array_equal(ref.generate(prompt, 16, use_graph=True), fus.generate(prompt, 16, use_graph=True))
    assert fus.launches_per_step() <= ref.launches_per_step() - 3 * cfg.n_layers
