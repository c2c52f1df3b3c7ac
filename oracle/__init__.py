"""CPU oracle for the quantized linear-layer hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may
import this package; nothing under ``blazr_b200/`` does.  PARITY UNPINNED (see quant_oracle.c header):
the reference holds no golden vector for this path and its arithmetic lives in crates whose source
is absent, so the oracle is pinned against gguf 0.19.0 ``gguf.quants`` (tests/golden) and follows
the reference loaders awq.rs / gptq.rs for the INT4 layouts.

Thin ctypes/numpy wrapper over ``libquant_oracle.so`` (built by ``make -C oracle``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libquant_oracle.so")

# ggml type ids (public spec)
Q4_0, Q4_1, Q5_0, Q5_1, Q8_0 = 2, 3, 6, 7, 8
Q2_K, Q3_K, Q4_K, Q5_K, Q6_K = 10, 11, 12, 13, 14
IQ4_NL, IQ4_XS = 20, 23
TQ1_0, TQ2_0 = 34, 35
IQ2_XXS, IQ2_XS, IQ3_XXS = 16, 17, 18
IQ3_S, IQ2_S = 21, 22
IQ1_S, IQ1_M = 19, 29
GGML_TYPES = {
    "Q4_0": Q4_0, "Q4_1": Q4_1, "Q5_0": Q5_0, "Q5_1": Q5_1, "Q8_0": Q8_0,
    "Q2_K": Q2_K, "Q3_K": Q3_K, "Q4_K": Q4_K, "Q5_K": Q5_K, "Q6_K": Q6_K,
    "IQ4_NL": IQ4_NL, "IQ4_XS": IQ4_XS, "TQ1_0": TQ1_0, "TQ2_0": TQ2_0,
    "IQ2_XXS": IQ2_XXS, "IQ2_XS": IQ2_XS, "IQ3_XXS": IQ3_XXS, "IQ2_S": IQ2_S, "IQ3_S": IQ3_S, "IQ1_S": IQ1_S, "IQ1_M": IQ1_M,
}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "clean"], check=True, capture_output=True)
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        try:
            if not os.path.exists(_LIB_PATH):
                build()
            _lib = C.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = C.CDLL(_LIB_PATH)
        _lib.orc_type_block_elems.restype = C.c_int64
        _lib.orc_type_block_bytes.restype = C.c_int64
        _lib.orc_type_sub.restype = C.c_int
        _lib.orc_gptq_perm.restype = C.c_int
    return _lib


def _p(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def block_elems(t: int) -> int:
    return int(lib().orc_type_block_elems(C.c_int(t)))


def block_bytes(t: int) -> int:
    return int(lib().orc_type_block_bytes(C.c_int(t)))


def sub_width(t: int) -> int:
    return int(lib().orc_type_sub(C.c_int(t)))


def row_bytes(t: int, K: int) -> int:
    return K // block_elems(t) * block_bytes(t)


def dequant_ggml(t: int, blocks: np.ndarray, N: int, K: int) -> np.ndarray:
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    out = np.empty((N, K), dtype=np.float32)
    lib().orc_dequant_ggml(C.c_int(t), _p(blocks), C.c_int64(N * K), _p(out))
    return out


def decompose_ggml(t: int, blocks: np.ndarray, N: int, K: int):
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    sub = sub_width(t)
    qi = np.empty((N, K), dtype=np.int8)
    a = np.empty((N, K // sub), dtype=np.float32)
    b = np.empty((N, K // sub), dtype=np.float32)
    lib().orc_decompose_ggml(C.c_int(t), _p(blocks), C.c_int64(N * K), _p(qi), _p(a), _p(b))
    return qi, a, b, sub


def quantize_act(x: np.ndarray):
    """x f32 [M,K] -> (q int8 [M,K], d f32 [M,K/32], bsum16 int32 [M,K/16])"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    M, K = x.shape
    q = np.empty((M, K), dtype=np.int8)
    d = np.empty((M, K // 32), dtype=np.float32)
    bs = np.empty((M, K // 16), dtype=np.int32)
    lib().orc_quantize_act(_p(x), C.c_int64(M), C.c_int64(K), _p(q), _p(d), _p(bs))
    return q, d, bs


def int_partials(qi: np.ndarray, xq: np.ndarray, sub: int) -> np.ndarray:
    N, K = qi.shape
    M = xq.shape[0]
    out = np.empty((M, N, K // sub), dtype=np.int32)
    lib().orc_int_partials(_p(qi), _p(xq), C.c_int(sub), C.c_int64(N), C.c_int64(K), C.c_int64(M), _p(out))
    return out


def matmul_dense(W: np.ndarray, X: np.ndarray, bias: np.ndarray | None = None) -> np.ndarray:
    """flavour A: Y[M,N] = X[M,K] @ W[N,K]^T (+bias), double accumulation"""
    W = np.ascontiguousarray(W, dtype=np.float32)
    X = np.ascontiguousarray(X, dtype=np.float32)
    N, K = W.shape
    M = X.shape[0]
    Y = np.empty((M, N), dtype=np.float32)
    lib().orc_matmul_dense(_p(W), _p(X), _p(bias), C.c_int64(N), C.c_int64(K), C.c_int64(M), _p(Y))
    return Y


def matmul_ggml_f32(t: int, blocks: np.ndarray, N: int, K: int, X: np.ndarray) -> np.ndarray:
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    X = np.ascontiguousarray(X, dtype=np.float32)
    M = X.shape[0]
    Y = np.empty((M, N), dtype=np.float32)
    lib().orc_matmul_ggml_f32(C.c_int(t), _p(blocks), C.c_int64(N), C.c_int64(K), _p(X), C.c_int64(M), _p(Y))
    return Y


def matmul_q8(qi, a, b, sub, X: np.ndarray, bias: np.ndarray | None = None) -> np.ndarray:
    """flavour B: int8 activations (per-32 scale), integer dots, double accumulation of the scaled partials"""
    N, K = qi.shape
    xq, xd, xbs = quantize_act(X)
    M = xq.shape[0]
    Y = np.empty((M, N), dtype=np.float32)
    lib().orc_matmul_q8(_p(qi), _p(a), _p(b), C.c_int(sub), C.c_int64(N), C.c_int64(K), _p(xq), _p(xd), _p(xbs),
                        C.c_int64(M), _p(bias), _p(Y))
    return Y


def matvec_ggml_q8(t: int, blocks: np.ndarray, N: int, K: int, xq, xd, xbs) -> np.ndarray:
    """packed-block flavour B (the timed CPU baseline kernel)"""
    M = xq.shape[0]
    Y = np.empty((M, N), dtype=np.float32)
    lib().orc_matvec_ggml_q8(C.c_int(t), _p(blocks), C.c_int64(N), C.c_int64(K), _p(xq), _p(xd), _p(xbs),
                             C.c_int64(M), _p(Y))
    return Y


def matvec_ggml_q8_fast(t: int, blocks: np.ndarray, N: int, K: int, xq, xd, xbs) -> np.ndarray:
    """the TIMED CPU baseline (cpu_fast.c): AVX2 + OpenMP packed-block kernels for Q8_0 / Q4_K / Q6_K, generic port otherwise"""
    M = xq.shape[0]
    Y = np.empty((M, N), dtype=np.float32)
    rc = lib().orc_matvec_fast(C.c_int(t), _p(blocks), C.c_int64(N), C.c_int64(K), _p(xq), _p(xd), _p(xbs), C.c_int64(M), _p(Y))
    if rc != 0:
        return matvec_ggml_q8(t, blocks, N, K, xq, xd, xbs)
    return Y


# ------------------------------------------------------------------------------------------------
# AWQ / GPTQ
# ------------------------------------------------------------------------------------------------
def awq_unpack_zeros(packed: np.ndarray, N: int) -> np.ndarray:
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    G = packed.shape[0]
    out = np.empty((G, N), dtype=np.float32)
    lib().orc_awq_unpack_zeros(_p(packed), C.c_int64(G), C.c_int64(N), _p(out))
    return out


def awq_dequant(qweight, scales, zeros, gs: int) -> np.ndarray:
    qweight = np.ascontiguousarray(qweight, dtype=np.uint32)
    scales = np.ascontiguousarray(scales, dtype=np.float32)
    zeros = np.ascontiguousarray(zeros, dtype=np.float32)
    K, n8 = qweight.shape
    N = n8 * 8
    out = np.empty((N, K), dtype=np.float32)
    lib().orc_awq_dequant(_p(qweight), _p(scales), _p(zeros), C.c_int64(gs), C.c_int64(N), C.c_int64(K), _p(out))
    return out


def awq_decompose(qweight, scales, zeros, gs: int):
    qweight = np.ascontiguousarray(qweight, dtype=np.uint32)
    scales = np.ascontiguousarray(scales, dtype=np.float32)
    zeros = np.ascontiguousarray(zeros, dtype=np.float32)
    K, n8 = qweight.shape
    N = n8 * 8
    qi = np.empty((N, K), dtype=np.int8)
    a = np.empty((N, K // 32), dtype=np.float32)
    b = np.empty((N, K // 32), dtype=np.float32)
    lib().orc_awq_decompose(_p(qweight), _p(scales), _p(zeros), C.c_int64(gs), C.c_int64(N), C.c_int64(K),
                            _p(qi), _p(a), _p(b))
    return qi, a, b, 32


def gptq_dequant(qweight, scales, qzeros, g_idx, gs: int, zero_plus_one: int = 1) -> np.ndarray:
    qweight = np.ascontiguousarray(qweight, dtype=np.uint32)
    scales = np.ascontiguousarray(scales, dtype=np.float32)
    qzeros = np.ascontiguousarray(qzeros, dtype=np.uint32)
    k8, N = qweight.shape
    K = k8 * 8
    gi = None if g_idx is None else np.ascontiguousarray(g_idx, dtype=np.int32)
    out = np.empty((N, K), dtype=np.float32)
    lib().orc_gptq_dequant(_p(qweight), _p(scales), _p(qzeros), _p(gi), C.c_int64(gs), C.c_int(zero_plus_one),
                           C.c_int64(N), C.c_int64(K), _p(out))
    return out


def gptq_perm(g_idx: np.ndarray, gs: int) -> np.ndarray:
    g_idx = np.ascontiguousarray(g_idx, dtype=np.int32)
    K = g_idx.shape[0]
    perm = np.empty(K, dtype=np.int32)
    rc = lib().orc_gptq_perm(_p(g_idx), C.c_int64(gs), C.c_int64(K), _p(perm))
    if rc != 0:
        raise ValueError("g_idx does not describe groups of exactly group_size members")
    return perm


def gptq_decompose(qweight, scales, qzeros, g_idx, gs: int, zero_plus_one: int = 1):
    """Returns (qi, a, b, 32, perm): decomposition in permuted-K order (perm None when g_idx is None)."""
    qweight = np.ascontiguousarray(qweight, dtype=np.uint32)
    scales = np.ascontiguousarray(scales, dtype=np.float32)
    qzeros = np.ascontiguousarray(qzeros, dtype=np.uint32)
    k8, N = qweight.shape
    K = k8 * 8
    perm = None if g_idx is None else gptq_perm(g_idx, gs)
    qi = np.empty((N, K), dtype=np.int8)
    a = np.empty((N, K // 32), dtype=np.float32)
    b = np.empty((N, K // 32), dtype=np.float32)
    lib().orc_gptq_decompose(_p(qweight), _p(scales), _p(qzeros), _p(perm), C.c_int64(gs), C.c_int(zero_plus_one),
                             C.c_int64(N), C.c_int64(K), _p(qi), _p(a), _p(b))
    return qi, a, b, 32, perm


def shard_range(total: int, rank: int, world: int):
    s, e = C.c_int64(), C.c_int64()
    lib().orc_shard_range(C.c_int64(total), C.c_int64(rank), C.c_int64(world), C.byref(s), C.byref(e))
    return int(s.value), int(e.value)


def det_exp(x: np.ndarray) -> np.ndarray:
    """deterministic f32 exp shared bit-for-bit with the GPU glue kernels (common.cuh det_expf)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().orc_det_expf(_p(x), C.c_int64(x.size), _p(out))
    return out
