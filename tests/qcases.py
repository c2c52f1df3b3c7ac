"""Shared builders for the GPU parity tests: one synthetic weight -> (device handle, oracle views)."""
from __future__ import annotations

import numpy as np

import oracle
from blazr_b200 import ops, synth

GGML_GPU = ["Q4_K", "Q6_K", "Q8_0", "Q4_0", "Q5_0", "IQ4_NL"]  # the last three run as exact re-encodings (formats.cuh adaptors)
INT4 = ["AWQ", "GPTQ", "GPTQ_ACT", "GPTQ_Z0", "AWQ_G64"]
ALL = GGML_GPU + INT4


class Case:
    """Holds the device weight and everything the oracle needs for the same weight."""

    def __init__(self, client, fmt: str, N: int, K: int, seed: int = 0, bias: bool = False):
        self.fmt, self.N, self.K = fmt, N, K
        self.perm = None
        self.bias = None
        if fmt in synth.GGML:
            t = synth.GGML[fmt]
            self.blocks = synth.random_ggml(t, N, K, seed=seed)
            self.t = t
            self.w = client.weight_from_ggml(t, self.blocks, N, K)
            self.qi, self.a, self.b, self.sub = oracle.decompose_ggml(t, self.blocks, N, K)
            self.deq = oracle.dequant_ggml(t, self.blocks, N, K)
        elif fmt.startswith("AWQ"):
            gs = 64 if fmt == "AWQ_G64" else 128
            qw, sc, zr = synth.random_awq(N, K, gs, seed=seed)
            self.src = (qw, sc, zr, gs)
            dq = ops.DecomposedQuantTensor(qw, sc, zr, None, ops.DecomposedQuantMethod("awq", gs), (N, K))
            self.w = client.weight_from_decomposed(dq)
            self.qi, self.a, self.b, self.sub = oracle.awq_decompose(qw, sc, zr, gs)
            self.deq = oracle.awq_dequant(qw, sc, zr, gs)
        elif fmt.startswith("GPTQ"):
            gs = 128
            act = fmt == "GPTQ_ACT"
            zpo = 0 if fmt == "GPTQ_Z0" else 1
            qw, sc, qz, gi, bs = synth.random_gptq(N, K, gs, seed=seed, act_order=act, bias=bias)
            self.bias = bs
            dq = ops.DecomposedQuantTensor(qw, sc, qz, gi, ops.DecomposedQuantMethod("gptq", gs), (N, K), bias=bs, zero_plus_one=zpo)
            self.w = client.weight_from_decomposed(dq)
            self.qi, self.a, self.b, self.sub, self.perm = oracle.gptq_decompose(qw, sc, qz, gi if act else None, gs, zpo)
            self.deq = oracle.gptq_dequant(qw, sc, qz, gi, gs, zpo)
        else:
            raise ValueError(fmt)

    def x_perm(self, x: np.ndarray) -> np.ndarray:
        """activations in the K order the decomposition uses (GPTQ act-order permutes K)"""
        return x if self.perm is None else np.ascontiguousarray(x[:, self.perm])

    def oracle_a(self, x: np.ndarray) -> np.ndarray:
        return oracle.matmul_dense(self.deq, x, self.bias)

    def oracle_b(self, x: np.ndarray) -> np.ndarray:
        return oracle.matmul_q8(self.qi, self.a, self.b, self.sub, self.x_perm(x), self.bias)
