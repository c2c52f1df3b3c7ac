"""Synthetic packed weights / activations.

There are no checkpoints in this environment, so every test and bench builds *packed bytes directly*:
quant payload bytes are uniform u8 (every bit pattern is a valid block), and the f16 super-block
scales ``d`` / ``dmin`` are drawn as U(0.5, 1.5) * sigma so that Var(W) ~= 1/K, which keeps activations
O(1) through a deep random-init model.  Pure numpy; no oracle, no GPU.
"""
from __future__ import annotations

import numpy as np

# ggml type ids (public ggml spec)
GGML = {
    "Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8,
    "Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14,
    "IQ4_NL": 20, "IQ4_XS": 23, "TQ1_0": 34, "TQ2_0": 35,
    "IQ2_XXS": 16, "IQ2_XS": 17, "IQ3_XXS": 18, "IQ3_S": 21, "IQ2_S": 22, "IQ1_S": 19, "IQ1_M": 29,
}
GGML_NAME = {v: k for k, v in GGML.items()}
# (block elems, block bytes)
GGML_SIZES = {
    2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34),
    10: (256, 84), 11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210),
    20: (32, 18), 23: (256, 136), 34: (256, 54), 35: (256, 66),
    16: (256, 66), 17: (256, 74), 18: (256, 98), 21: (256, 110), 22: (256, 82), 19: (256, 50), 29: (256, 56),
}
# per type: (byte offsets of f16 d fields, byte offsets of f16 min fields, min/d ratio that zeroes the
# mean of W, std of W at d = 1 with that ratio) -- measured once with uniform payload bytes.
_FIELDS = {
    2: ([0], [], 0.0, 4.61), 3: ([0], [2], -7.5, 4.60), 6: ([0], [], 0.0, 9.24), 7: ([0], [2], -15.5, 9.23),
    8: ([0], [], 0.0, 74.0), 10: ([80], [82], 1.5, 13.9), 11: ([108], [], 0.0, 43.3),
    12: ([0], [2], 7.5, 259.2), 13: ([0], [2], 15.55, 527.5), 14: ([208], [], 0.0, 1363.9),
    20: ([0], [], 0.0, 67.3), 23: ([0], [], 0.0, 1249.0), 34: ([52], [], 0.0, 0.82), 35: ([64], [], 0.0, 1.22),
    16: ([0], [], 0.0, 53.1), 17: ([0], [], 0.0, 55.8), 18: ([0], [], 0.0, 142.7),
    21: ([0], [], 0.0, 140.3), 22: ([0], [], 0.0, 56.5), 19: ([0], [], 0.0, 7.21),
    29: ([], [], 0.0, 7.21),  # IQ1_M: the f16 d is scattered over the top nibbles of its four u16 scale words (set_iq1m_d)
}


def ggml_row_bytes(t: int, K: int) -> int:
    be, bb = GGML_SIZES[t]
    assert K % be == 0, f"K={K} not a multiple of the {GGML_NAME[t]} block ({be})"
    return K // be * bb


def ggml_bytes_per_weight(t: int) -> float:
    be, bb = GGML_SIZES[t]
    return bb / be


def set_iq1m_d(blk: np.ndarray, d: np.ndarray) -> None:
    """IQ1_M keeps its f16 super-block scale in the top nibbles of the four u16 scale words (bytes 48..55):
    d = s0 >> 12 | (s1 >> 8) & 0xF0 | (s2 >> 4) & 0xF00 | s3 & 0xF000.  blk: uint8 [nb, 56] (modified in place)."""
    bits = np.ascontiguousarray(d, dtype=np.float16).view(np.uint16).astype(np.uint16)
    sc = blk[:, 48:56].copy().view(np.uint16)
    for i, sh in enumerate((0, 4, 8, 12)):
        sc[:, i] = (sc[:, i] & np.uint16(0x0FFF)) | (((bits >> np.uint16(sh)) & np.uint16(0xF)) << np.uint16(12))
    blk[:, 48:56] = sc.view(np.uint8)


def random_ggml(t: int, N: int, K: int, seed: int = 0, gain: float = 1.0) -> np.ndarray:
    """Random raw ggml blocks for a logical [N, K] weight (row-major rows of K/block blocks)."""
    be, bb = GGML_SIZES[t]
    nb = N * (K // be)
    rng = np.random.Generator(np.random.PCG64(seed))
    blk = rng.integers(0, 256, size=(nb, bb), dtype=np.uint8)
    d_off, m_off, ratio, std1 = _FIELDS[t]
    sigma = gain / (std1 * np.sqrt(K))
    d = (rng.uniform(0.5, 1.5, size=nb) * sigma).astype(np.float16)
    for o in d_off:
        blk[:, o:o + 2] = d.view(np.uint8).reshape(nb, 2)
    if t == 29:
        set_iq1m_d(blk, d)
    if m_off:
        m = (d.astype(np.float32) * ratio).astype(np.float16)
        for o in m_off:
            blk[:, o:o + 2] = m.view(np.uint8).reshape(nb, 2)
    return blk.reshape(N, K // be * bb)


def random_awq(N: int, K: int, gs: int = 128, seed: int = 0, gain: float = 1.0):
    """AWQ triplet exactly as blazr hands it over (reference src/loader/safetensors/awq.rs:190-226):
    qweight u32 [K, N/8] (AWQ nibble order), scales f32 [K/gs, N] (f16-representable),
    zeros f32 [K/gs, N] (already unpacked integers 0..15)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qweight = rng.integers(0, 2 ** 32, size=(K, N // 8), dtype=np.uint64).astype(np.uint32)
    G = K // gs
    scales = (rng.uniform(0.5, 1.5, size=(G, N)) * gain / (4.6 * np.sqrt(K))).astype(np.float16).astype(np.float32)
    zeros = rng.integers(0, 16, size=(G, N)).astype(np.float32)
    return qweight, scales, zeros


def random_gptq(N: int, K: int, gs: int = 128, seed: int = 0, act_order: bool = False, bias: bool = False,
                gain: float = 1.0):
    """GPTQ group as blazr hands it over (reference src/loader/safetensors/gptq.rs:198-259):
    qweight u32 [K/8, N], scales f32 [G, N], qzeros u32 [G, N/8] (packed), g_idx i32 [K] | None,
    bias f32 [N] | None."""
    rng = np.random.Generator(np.random.PCG64(seed))
    qweight = rng.integers(0, 2 ** 32, size=(K // 8, N), dtype=np.uint64).astype(np.uint32)
    G = K // gs
    scales = (rng.uniform(0.5, 1.5, size=(G, N)) * gain / (4.6 * np.sqrt(K))).astype(np.float16).astype(np.float32)
    qzeros = rng.integers(0, 2 ** 32, size=(G, N // 8), dtype=np.uint64).astype(np.uint32)
    if act_order:
        g_idx = np.repeat(np.arange(G, dtype=np.int32), gs)
        g_idx = g_idx[rng.permutation(K)].astype(np.int32)
    else:
        g_idx = (np.arange(K) // gs).astype(np.int32)
    b = (rng.standard_normal(N) * 0.1).astype(np.float16).astype(np.float32) if bias else None
    return qweight, scales, qzeros, g_idx, b


def random_act(M: int, K: int, seed: int = 7, dtype=np.float32) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((M, K)).astype(np.float32).astype(dtype)
