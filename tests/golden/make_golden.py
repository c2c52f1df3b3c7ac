"""Generates tests/golden/ggml_dequant.npz -- the fixture that pins the CPU oracle.

The reference (blazr) holds no numeric golden vector for the quantized-matmul path and its arithmetic
lives in un-vendored crates, so the oracle is pinned against the independent numpy
implementation of the public ggml formats that ships in this image: gguf 0.19.0 ``gguf.quants``.
For every format: 64 random packed blocks (payload bytes uniform, f16 scales finite) and the f32
values ``gguf.quants.dequantize`` returns for them.  AWQ nibble-order golden: AutoAWQ's documented
``order_map = [0,2,4,6,1,3,5,7]`` packing round-tripped through blazr's AWQ_SHIFTS
(reference src/loader/safetensors/awq.rs:29-32).

Run here (gguf importable):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from blazr_b200 import synth  # noqa: E402

import gguf  # noqa: E402
from gguf import quants  # noqa: E402


def main():
    out = {}
    for name, t in synth.GGML.items():
        be, bb = synth.GGML_SIZES[t]
        n_blocks = 64
        K = be * 8
        N = n_blocks // 8
        blocks = synth.random_ggml(t, N, K, seed=0x601D + t, gain=50.0)
        # also exercise extreme payloads: first block all 0x00, second all 0xFF payload (scales kept finite)
        d_off, m_off, _, _ = synth._FIELDS[t]
        flat = blocks.reshape(n_blocks, bb).copy()
        keep = set()
        for o in d_off + m_off:
            keep.update((o, o + 1))
        for blk_i, fill in ((0, 0x00), (1, 0xFF)):
            for j in range(bb):
                if j not in keep:
                    flat[blk_i, j] = fill
        if t == 29:  # IQ1_M: restore the scattered f16 scale of the two extreme-payload blocks
            keep_d = np.array([0.02, 0.03], dtype=np.float16)
            synth.set_iq1m_d(flat[:2], keep_d)
        qt = gguf.GGMLQuantizationType(t)
        deq = quants.dequantize(flat.reshape(N, -1), qt).astype(np.float32)
        assert deq.shape == (N, K)
        out[f"{name}_blocks"] = flat.reshape(N, -1)
        out[f"{name}_deq"] = deq
    # Q8_0 quantizer golden ("bit-exact same results as reference implementation in ggml-quants.c")
    x = synth.random_act(4, 256, seed=99)
    x[0, :32] = 0.0  # an all-zero block
    out["q8_0_quant_x"] = x
    out["q8_0_quant_blocks"] = quants.quantize(x, gguf.GGMLQuantizationType.Q8_0)
    # AWQ order map golden: pack values v[j] (j = logical column 0..7) the AutoAWQ way
    order_map = [0, 2, 4, 6, 1, 3, 5, 7]
    rng = np.random.default_rng(5)
    vals = rng.integers(0, 16, size=(16, 8), dtype=np.uint32)
    packed = np.zeros(16, dtype=np.uint32)
    for i in range(8):
        packed |= vals[:, order_map[i]] << np.uint32(4 * i)
    out["awq_order_vals"] = vals
    out["awq_order_packed"] = packed
    path = os.path.join(os.path.dirname(__file__), "ggml_dequant.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
