"""GPU parity of the grouped expert-bank matvec (b200q_moe_matmul_q8) and the decode MoE MLP against the CPU
oracle, on DeepSeek-V2-Lite-shaped experts (gate/up 1408 x 2048, down 2048 x 1408: K is not a multiple of
256, so down is a 32-block format -- SURVEY.md section 7 "format corner cases").  Every slot must equal the
plain per-weight matvec bit for bit (same kernel arithmetic) and the oracle's int8 flavour."""
import numpy as np
import pytest
import torch

import oracle
from blazr_b200 import ops, synth, tp
from test_gpu_quant import assert_bit_exact, rel_err

pytestmark = pytest.mark.gpu


def _bank(client, fmt, E, N, K, seed0):
    t = synth.GGML[fmt]
    blocks = [synth.random_ggml(t, N, K, seed=seed0 + e) for e in range(E)]
    ws = [client.weight_from_ggml(t, b, N, K) for b in blocks]
    dec = [oracle.decompose_ggml(t, b, N, K) for b in blocks]
    return blocks, ws, dec


@pytest.mark.parametrize("fmt,N,K", [("Q4_K", 2816, 2048), ("Q6_K", 256, 512), ("Q8_0", 2048, 1408), ("Q8_0", 130, 96)])
@pytest.mark.parametrize("shared_x", [True, False])
def test_grouped_matvec_matches_oracle_and_single(client, fmt, N, K, shared_x):
    E, n_slots = 6, 8
    _, ws, dec = _bank(client, fmt, E, N, K, 100)
    bank = ops.ExpertBank(ws)
    sel_h = np.array([3, 0, 5, 5, 1, 2, 4, 0], dtype=np.int32)
    sel = torch.from_numpy(sel_h).cuda()
    rows = 2 if shared_x else n_slots
    div = 4 if shared_x else 1
    x = synth.random_act(rows, K, seed=9)
    xq = client.quantize_act(torch.from_numpy(x).cuda())
    y = bank.matmul_q8(sel, xq, rows, div).cpu().numpy()
    for s in range(n_slots):
        qi, a, b, sub = dec[sel_h[s]]
        xr = x[s // div:s // div + 1]
        ref = oracle.matmul_q8(qi, a, b, sub, xr)
        assert_bit_exact(y[s:s + 1], ref, f"slot {s}")
        single = client.matmul_q8(client.quantize_act(torch.from_numpy(xr).cuda()), 1, ws[sel_h[s]]).cpu().numpy()
        assert np.array_equal(single.view(np.uint32), y[s:s + 1].view(np.uint32))


def test_grouped_matvec_skips_negative_slots(client):
    E, N, K = 4, 384, 512
    _, ws, dec = _bank(client, "Q4_K", E, N, K, 200)
    bank = ops.ExpertBank(ws)
    sel_h = np.array([-1, 2, -1, 0, 3, -1], dtype=np.int32)
    x = synth.random_act(1, K, seed=3)
    xq = client.quantize_act(torch.from_numpy(x).cuda())
    y = bank.matmul_q8(torch.from_numpy(sel_h).cuda(), xq, 1, len(sel_h)).cpu().numpy()
    for s, e in enumerate(sel_h):
        if e < 0:
            assert not y[s].any()
        else:
            qi, a, b, sub = dec[e]
            assert_bit_exact(y[s:s + 1], oracle.matmul_q8(qi, a, b, sub, x), f"slot {s}")


def test_bank_set_swaps_an_expert(client):
    """set_expert_weights analogue (reference executor_cache.rs:283): the pointer table entry changes, nothing else"""
    E, N, K = 3, 256, 256
    _, ws, dec = _bank(client, "Q8_0", E, N, K, 300)
    bank = ops.ExpertBank(ws)
    t = synth.GGML["Q8_0"]
    nb = synth.random_ggml(t, N, K, seed=999)
    nw = client.weight_from_ggml(t, nb, N, K)
    bank.set(1, nw)
    torch.cuda.synchronize()
    x = synth.random_act(1, K, seed=4)
    xq = client.quantize_act(torch.from_numpy(x).cuda())
    y = bank.matmul_q8(torch.tensor([1], dtype=torch.int32, device="cuda"), xq, 1, 1).cpu().numpy()
    qi, a, b, sub = oracle.decompose_ggml(t, nb, N, K)
    assert_bit_exact(y, oracle.matmul_q8(qi, a, b, sub, x), "swapped expert")
    with pytest.raises(ops.B200QError):
        bank.set(0, client.weight_from_ggml(t, synth.random_ggml(t, 128, K, seed=1), 128, K))


def _oracle_moe(x, sel, gw, gu_dec, dn_dec, ffn):
    T, top_k = sel.shape
    out = np.zeros((T, x.shape[1]), dtype=np.float32)
    ys = []
    for tkn in range(T):
        acc = None
        for j in range(top_k):
            e = sel[tkn, j]
            if e < 0:
                y = np.zeros((1, x.shape[1]), dtype=np.float32)
            else:
                gu = oracle.matmul_q8(*gu_dec[e], x[tkn:tkn + 1])
                g, u = gu[:, :ffn], gu[:, ffn:]
                act = ((g / (np.float32(1.0) + oracle.det_exp(-g))).astype(np.float32) * u).astype(np.float32)
                if ffn % 256:
                    act = np.concatenate([act, np.zeros((1, 256 - ffn % 256), dtype=np.float32)], axis=1)
                qi, a, b, sub = dn_dec[e]
                Kp = act.shape[1]
                qi_p = np.zeros((qi.shape[0], Kp), dtype=qi.dtype); qi_p[:, :qi.shape[1]] = qi
                a_p = np.zeros((a.shape[0], Kp // sub), dtype=a.dtype); a_p[:, :a.shape[1]] = a
                b_p = np.zeros((b.shape[0], Kp // sub), dtype=b.dtype); b_p[:, :b.shape[1]] = b
                y = oracle.matmul_q8(qi_p, a_p, b_p, sub, act)
            ys.append(y)
    return np.concatenate(ys, axis=0)


def test_moe_mlp_decode_deepseek_v2_lite_shape(client):
    """64 routed experts are too slow for the oracle; 8 experts of the real shape, top-3, 2 tokens"""
    hidden, ffn, E, top_k, T = 2048, 1408, 8, 3, 2
    _, gu_w, gu_dec = _bank(client, "Q4_K", E, 2 * ffn, hidden, 400)
    _, dn_w, dn_dec = _bank(client, "Q8_0", E, hidden, ffn, 500)
    moe = ops.MoeMlp(client, [ops.ExpertWeights(g, d) for g, d in zip(gu_w, dn_w)], ffn, hidden)
    assert moe.get_expert_weights(3).down_proj is dn_w[3]
    x = synth.random_act(T, hidden, seed=21)
    sel_h = np.array([[1, 6, 3], [0, 6, 7]], dtype=np.int32)
    gw_h = np.array([[0.5, 0.3, 0.2], [0.6, 0.25, 0.15]], dtype=np.float32)
    xt = torch.from_numpy(x).cuda()
    out = moe.forward_decode(xt, torch.from_numpy(sel_h).cuda(), torch.from_numpy(gw_h).cuda()).cpu().numpy()
    ys = _oracle_moe(x, sel_h, gw_h, gu_dec, dn_dec, ffn)                    # [T*top_k, hidden], oracle per slot
    ref = (torch.from_numpy(ys).reshape(T, top_k, hidden) * torch.from_numpy(gw_h).unsqueeze(-1)).sum(dim=1).numpy()
    assert rel_err(out, ref) < 1e-6
    # expert parallelism, 2 ranks emulated on one GPU: masked local slots, partial outputs summed
    total = np.zeros_like(out)
    for rank in range(2):
        e0, e1 = tp.expert_range(E, rank, 2)
        local = ops.MoeMlp(client, [ops.ExpertWeights(g, d) for g, d in zip(gu_w[e0:e1], dn_w[e0:e1])], ffn, hidden)
        ls, lg = tp.ep_local_slots(torch.from_numpy(sel_h).cuda(), torch.from_numpy(gw_h).cuda(), e0, e1)
        total += local.forward_decode(xt, ls, lg).cpu().numpy()
    assert rel_err(total, ref) < 1e-6


def test_moe_interleaved_banks_with_swiglu_epilogue_equal_plain_banks(client):
    """gate|up experts uploaded in the SwiGLU-epilogue row order: the grouped gate|up launch activates and quantises itself
    (4 launches per MoE layer instead of 5) -- same output bits as the plain banks; also under CUDA-graph replay with the
    selection changing between replays, and with a slot masked to -1 (expert hosted by another rank)"""
    hidden, ffn, E, top_k, T = 2048, 1408, 8, 3, 2
    t4, t8 = synth.GGML["Q4_K"], synth.GGML["Q8_0"]
    gub = [synth.random_ggml(t4, 2 * ffn, hidden, seed=600 + e) for e in range(E)]
    dnb = [synth.random_ggml(t8, hidden, ffn, seed=700 + e) for e in range(E)]
    order = ops.gate_up_row_order(ffn)
    dn = [client.weight_from_ggml(t8, b_, hidden, ffn) for b_ in dnb]
    plain = ops.MoeMlp(client, [ops.ExpertWeights(client.weight_from_ggml(t4, b_, 2 * ffn, hidden), d) for b_, d in zip(gub, dn)], ffn, hidden)
    il = ops.MoeMlp(client, [ops.ExpertWeights(client.weight_from_ggml(t4, np.ascontiguousarray(b_[order]), 2 * ffn, hidden), d, interleaved=True)
                             for b_, d in zip(gub, dn)], ffn, hidden)
    assert il.interleaved and not plain.interleaved
    x = torch.from_numpy(synth.random_act(T, hidden, seed=33)).cuda()
    sel = torch.tensor([[1, 6, 3], [0, -1, 7]], dtype=torch.int32, device="cuda")
    gw = torch.tensor([[0.5, 0.3, 0.2], [0.6, 0.0, 0.4]], dtype=torch.float32, device="cuda")
    n0 = ops.launch_count()
    a = plain.forward_decode(x, sel, gw).clone()
    n1 = ops.launch_count()
    b = il.forward_decode(x, sel, gw).clone()
    n2 = ops.launch_count()
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert (n1 - n0, n2 - n1) == (5, 4)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            out = il.forward_decode(x, sel, gw)
    for new_sel in ([[2, 5, 4], [7, 0, 1]], [[1, 6, 3], [0, -1, 7]]):
        sel.copy_(torch.tensor(new_sel, dtype=torch.int32))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, plain.forward_decode(x, sel, gw))
