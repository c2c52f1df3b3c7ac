"""CPU oracle of one Llama-family decode step -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Mirrors the caller pattern of the reference (src/engine/executor_generate.rs:362-405: sample -> forward ->
read token) with the oracle's int8-activation arithmetic ("flavour B") for every quantized projection, f32
everywhere else.  PARITY UNPINNED: boostr's forward is not vendored; this restates the public Llama
architecture (RMSNorm, adjacent-pair RoPE as in ggml, GQA attention, SwiGLU) and greedy argmax with the
lowest index winning ties (reference src/engine/executor_cache.rs:189-196 takes the last position's argmax).
"""
from __future__ import annotations

import numpy as np

import oracle


class OracleLinear:
    def __init__(self, hl, flavour: str = "B"):
        """hl: blazr_b200.decode.HostLinear.  flavour "B" = int8 activations + integer dots (the decode contract),
        "A" = dequantized f32 weights x f32 activations in f64 (the reference's f32 CPU path; tolerance contract)"""
        self.N, self.K = hl.N, hl.K
        self.perm = None
        self.flavour = flavour
        if hl.fmt in oracle.GGML_TYPES:
            t = oracle.GGML_TYPES[hl.fmt]
            if flavour == "A":
                self.deq = oracle.dequant_ggml(t, hl.data, hl.N, hl.K)
            else:
                self.qi, self.a, self.b, self.sub = oracle.decompose_ggml(t, hl.data, hl.N, hl.K)
        elif hl.fmt == "AWQ":
            qw, sc, zr, gs = hl.data
            if flavour == "A":
                self.deq = oracle.awq_dequant(qw, sc, zr, gs)
            else:
                self.qi, self.a, self.b, self.sub = oracle.awq_decompose(qw, sc, zr, gs)
        elif hl.fmt == "GPTQ":
            qw, sc, qz, gi, gs = hl.data
            if flavour == "A":
                self.deq = oracle.gptq_dequant(qw, sc, qz, None, gs, 1)
            else:
                self.qi, self.a, self.b, self.sub, self.perm = oracle.gptq_decompose(qw, sc, qz, None, gs, 1)
        else:
            raise ValueError(hl.fmt)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        if self.flavour == "A":
            return oracle.matmul_dense(self.deq, x)
        return oracle.matmul_q8(self.qi, self.a, self.b, self.sub, x)


def rmsnorm(h: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    """f64 sum of squares (exact products, order independent), then the f32 operation sequence of the device
    kernel: inv = 1 / sqrt(ss / H + eps); y = (h * inv) * w"""
    h = h.astype(np.float32)
    ss = np.float32((h.astype(np.float64) ** 2).sum(axis=-1, keepdims=True))
    t = (ss / np.float32(h.shape[-1])).astype(np.float32) + np.float32(eps)
    inv = (np.float32(1.0) / np.sqrt(t.astype(np.float32))).astype(np.float32)
    return ((h * inv).astype(np.float32) * w).astype(np.float32)


def rope_table(max_ctx: int, head_dim: int, theta: float) -> np.ndarray:
    """[max_ctx, hd/2, 2] f32 (cos, sin) of pos * theta^(-2i/hd): f64 then one rounding (same formula as the
    harness uploads to the device, restated here so the oracle does not import product code)"""
    i = np.arange(head_dim // 2, dtype=np.float64)
    freq = np.power(float(theta), -2.0 * i / head_dim)
    ang = np.arange(max_ctx, dtype=np.float64)[:, None] * freq[None, :]
    return np.stack([np.cos(ang), np.sin(ang)], axis=-1).astype(np.float32)


def rope_pairs(x: np.ndarray, cs: np.ndarray) -> np.ndarray:
    """x [..., hd]; rotates adjacent pairs (2i, 2i+1); cs [hd/2, 2] = (cos, sin) of this position.
    Separate f32 multiplies and add/sub (no FMA), as the device kernel does."""
    c, s = cs[:, 0], cs[:, 1]
    x0, x1 = x[..., 0::2], x[..., 1::2]
    out = np.empty_like(x)
    out[..., 0::2] = (x0 * c).astype(np.float32) - (x1 * s).astype(np.float32)
    out[..., 1::2] = (x0 * s).astype(np.float32) + (x1 * c).astype(np.float32)
    return out


class OracleModel:
    def __init__(self, host, flavour: str = "B"):
        """host: blazr_b200.decode.HostModel"""
        self.cfg = host.cfg
        self.embed = host.embed
        self.layers = []
        for lay in host.layers:
            self.layers.append({k: (OracleLinear(v, flavour) if k in ("q", "k", "v", "o", "gate", "up", "down") else v) for k, v in lay.items()})
        self.final_norm = host.final_norm
        self.head = OracleLinear(host.lm_head, flavour)
        self.rope = rope_table(4096, self.cfg.head_dim, self.cfg.rope_theta)
        self.reset()

    def reset(self):
        self.k_cache = [[] for _ in self.layers]
        self.v_cache = [[] for _ in self.layers]
        self.pos = 0

    def step(self, token: int) -> np.ndarray:
        """one token in -> logits f32 [V]"""
        cfg = self.cfg
        h = self.embed[token].astype(np.float32)[None, :]
        nh, nkv, hd = cfg.n_heads, cfg.n_kv_heads, cfg.head_dim
        for li, lay in enumerate(self.layers):
            x = rmsnorm(h, lay["attn_norm"], cfg.eps)
            q = lay["q"](x).reshape(nh, hd)
            k = lay["k"](x).reshape(nkv, hd)
            v = lay["v"](x).reshape(nkv, hd)
            q = rope_pairs(q, self.rope[self.pos])
            k = rope_pairs(k, self.rope[self.pos])
            self.k_cache[li].append(k)
            self.v_cache[li].append(v)
            K = np.stack(self.k_cache[li], axis=0).astype(np.float64)  # [T, nkv, hd]
            V = np.stack(self.v_cache[li], axis=0).astype(np.float64)
            rep = nh // nkv
            out = np.empty((nh, hd), dtype=np.float32)
            scale = np.float32(1.0) / np.sqrt(np.float32(hd))
            for hh in range(nh):
                kv = hh // rep
                sc = (K[:, kv, :] @ q[hh].astype(np.float64)).astype(np.float32) * scale   # f64 dot, one rounding, f32 scale
                p = oracle.det_exp((sc - sc.max()).astype(np.float32))
                den = np.float32(p.astype(np.float64).sum())
                num = (p.astype(np.float64)[:, None] * V[:, kv, :]).sum(axis=0).astype(np.float32)
                out[hh] = num / den
            h = (h + lay["o"](out.reshape(1, nh * hd))).astype(np.float32)
            x = rmsnorm(h, lay["mlp_norm"], cfg.eps)
            g = lay["gate"](x)
            u = lay["up"](x)
            act = ((g / (np.float32(1.0) + oracle.det_exp(-g))).astype(np.float32) * u).astype(np.float32)
            h = (h + lay["down"](act)).astype(np.float32)
        x = rmsnorm(h, self.final_norm, cfg.eps)
        logits = self.head(x)[0]
        self.pos += 1
        return logits

    def generate(self, prompt, n_new: int):
        """greedy; returns (tokens [n_new], relative top-2 gap of every generated step [n_new])"""
        self.reset()
        toks, gaps = [], []
        logits = None
        for t in prompt:
            logits = self.step(int(t))
        for _ in range(n_new):
            nxt = int(np.argmax(logits))
            top2 = np.partition(logits, -2)[-2:]
            gap = float((top2[1] - top2[0]) / max(np.abs(logits).max(), 1e-30))
            gaps.append(gap)
            toks.append(nxt)
            logits = self.step(nxt)
        return np.asarray(toks, dtype=np.int64), np.asarray(gaps)
