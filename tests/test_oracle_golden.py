"""CPU: the oracle against its pins -- gguf.quants golden vectors (tests/golden/make_golden.py), the AWQ
nibble-order golden, and its own internal consistency (decomposition == dequant, flavour A ~ flavour B)."""
import os

import numpy as np
import pytest

import oracle
from blazr_b200 import synth

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ggml_dequant.npz"))


@pytest.mark.parametrize("name", list(synth.GGML))
def test_dequant_matches_gguf_quants_bit_exact(name):
    t = synth.GGML[name]
    blk, deq = GOLD[f"{name}_blocks"], GOLD[f"{name}_deq"]
    N, K = deq.shape
    mine = oracle.dequant_ggml(t, blk, N, K)
    assert np.array_equal(mine.view(np.uint32), deq.view(np.uint32))


@pytest.mark.parametrize("name", list(synth.GGML))
def test_decomposition_reproduces_dequant(name):
    t = synth.GGML[name]
    be, _ = synth.GGML_SIZES[t]
    N, K = 8, be * 8 if be == 256 else 256
    blk = synth.random_ggml(t, N, K, seed=3)
    qi, a, b, sub = oracle.decompose_ggml(t, blk, N, K)
    w = (a.repeat(sub, axis=1) * qi.astype(np.float32)) - b.repeat(sub, axis=1)
    assert np.array_equal(w.view(np.uint32), oracle.dequant_ggml(t, blk, N, K).view(np.uint32))


def test_block_sizes_match_synth_table():
    for name, t in synth.GGML.items():
        assert (oracle.block_elems(t), oracle.block_bytes(t)) == synth.GGML_SIZES[t], name


def test_q8_quantizer_matches_ggml_q8_0():
    x = GOLD["q8_0_quant_x"]
    qb = GOLD["q8_0_quant_blocks"].reshape(4, 8, 34)
    q, d, bs = oracle.quantize_act(x)
    dref = qb[:, :, :2].copy().view(np.float16).astype(np.float32).reshape(4, 8)
    qref = qb[:, :, 2:].copy().view(np.int8).reshape(4, 256)
    assert np.array_equal(q, qref)
    assert np.array_equal(d.astype(np.float16).astype(np.float32), dref)
    assert np.array_equal(bs.reshape(4, 8, 2).sum(-1), q.reshape(4, 8, 32).astype(np.int32).sum(-1))
    assert np.all(q[0, :32] == 0) and d[0, 0] == 0.0


def test_awq_shifts_roundtrip_autoawq_order():
    vals, packed = GOLD["awq_order_vals"], GOLD["awq_order_packed"]
    z = oracle.awq_unpack_zeros(packed.reshape(16, 1), 8)
    assert np.array_equal(z.astype(np.uint32), vals)


def test_awq_dequant_formula():
    N, K, gs = 16, 256, 128
    qw, sc, zr = synth.random_awq(N, K, gs, seed=1)
    w = oracle.awq_dequant(qw, sc, zr, gs)
    shifts = [0, 16, 4, 20, 8, 24, 12, 28]
    for n, k in [(0, 0), (5, 17), (15, 255), (9, 128)]:
        q = (int(qw[k, n // 8]) >> shifts[n % 8]) & 0xF
        assert w[n, k] == np.float32(np.float32(q) - zr[k // gs, n]) * sc[k // gs, n]
    qi, a, b, sub = oracle.awq_decompose(qw, sc, zr, gs)
    assert np.array_equal((a.repeat(sub, 1) * qi.astype(np.float32)).view(np.uint32), w.view(np.uint32))


@pytest.mark.parametrize("act_order", [False, True])
def test_gptq_dequant_and_perm(act_order):
    N, K, gs = 16, 512, 128
    qw, sc, qz, gi, _ = synth.random_gptq(N, K, gs, seed=2, act_order=act_order)
    w = oracle.gptq_dequant(qw, sc, qz, gi, gs, 1)
    for n, k in [(0, 0), (7, 300), (15, 511)]:
        q = (int(qw[k // 8, n]) >> (4 * (k % 8))) & 0xF
        g = int(gi[k])
        z = ((int(qz[g, n // 8]) >> (4 * (n % 8))) & 0xF) + 1
        assert w[n, k] == np.float32(q - z) * sc[g, n]
    qi, a, b, sub, perm = oracle.gptq_decompose(qw, sc, qz, gi if act_order else None, gs, 1)
    wp = a.repeat(sub, 1) * qi.astype(np.float32)
    if act_order:
        assert np.array_equal(np.sort(perm), np.arange(K))
        assert np.all(np.diff(gi[perm]) >= 0)
        assert np.array_equal(wp.view(np.uint32), w[:, perm].view(np.uint32))
    else:
        assert np.array_equal(wp.view(np.uint32), w.view(np.uint32))


def test_gptq_perm_rejects_ragged_groups():
    gi = np.zeros(256, dtype=np.int32)
    with pytest.raises(ValueError):
        oracle.gptq_perm(gi, 128)


@pytest.mark.parametrize("name", ["Q8_0", "Q4_K", "Q6_K", "Q5_K", "Q2_K", "Q3_K", "Q4_0", "IQ4_XS"])
def test_flavours_agree(name):
    """int8-activation flavour B stays within the north_star tolerance (1e-2 rel) of the f32 flavour A"""
    t = synth.GGML[name]
    N, K, M = 64, 1024, 3
    blk = synth.random_ggml(t, N, K, seed=11)
    x = synth.random_act(M, K)
    ya = oracle.matmul_ggml_f32(t, blk, N, K, x)
    ya2 = oracle.matmul_dense(oracle.dequant_ggml(t, blk, N, K), x)
    assert np.allclose(ya, ya2, rtol=0, atol=1e-6 * np.abs(ya).max())
    qi, a, b, sub = oracle.decompose_ggml(t, blk, N, K)
    yb = oracle.matmul_q8(qi, a, b, sub, x)
    assert np.abs(ya - yb).max() <= 1e-2 * np.abs(ya).max()
    xq, xd, xbs = oracle.quantize_act(x)
    yc = oracle.matvec_ggml_q8(t, blk, N, K, xq, xd, xbs)
    assert np.abs(yc - yb).max() <= 1e-4 * np.abs(yb).max()


def test_int_partials_sum_to_flavour_b():
    t = synth.GGML["Q6_K"]
    N, K = 8, 512
    blk = synth.random_ggml(t, N, K, seed=5)
    x = synth.random_act(2, K)
    qi, a, b, sub = oracle.decompose_ggml(t, blk, N, K)
    xq, xd, xbs = oracle.quantize_act(x)
    part = oracle.int_partials(qi, xq, sub)
    ref = (qi.astype(np.int32).reshape(1, N, K // sub, sub) * xq.astype(np.int32).reshape(2, 1, K // sub, sub)).sum(-1)
    assert np.array_equal(part, ref)


def test_shard_range_reference_rule():
    # reference src/engine/tensor_parallel.rs:169-185 checks exactly these splits
    assert oracle.shard_range(32, 0, 4) == (0, 8) and oracle.shard_range(32, 3, 4) == (24, 32)  # test_shard_range_even
    assert [oracle.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]  # test_shard_range_uneven
    assert [oracle.shard_range(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            spans = [oracle.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("name", ["Q8_0", "Q4_K", "Q6_K"])
@pytest.mark.parametrize("M", [1, 3])
def test_timed_cpu_baseline_matches_exact_oracle(name, M):
    """oracle/cpu_fast.c (the AVX2 kernels bench.py times as the CPU baseline) computes the same int8-activation matvec as
    the exact-accumulation oracle, to f32 accumulation accuracy"""
    from blazr_b200 import synth
    t = synth.GGML[name]
    N, K = 300, 1024
    blk = synth.random_ggml(t, N, K, seed=4)
    x = synth.random_act(M, K, seed=5)
    xq, xd, xbs = oracle.quantize_act(x)
    y = oracle.matvec_ggml_q8_fast(t, blk, N, K, xq, xd, xbs)
    qi, a, b, sub = oracle.decompose_ggml(t, blk, N, K)
    ref = oracle.matmul_q8(qi, a, b, sub, x)
    assert float(np.abs(y - ref).max() / np.abs(ref).max()) < 1e-5
    # a type without a fast kernel falls back to the generic port
    t2 = synth.GGML["Q5_K"]
    blk2 = synth.random_ggml(t2, 64, 512, seed=6)
    x2 = synth.random_act(1, 512, seed=7)
    q2 = oracle.quantize_act(x2)
    assert np.array_equal(oracle.matvec_ggml_q8_fast(t2, blk2, 64, 512, *q2), oracle.matvec_ggml_q8(t2, blk2, 64, 512, *q2))
