"""CPU: the C-ABI shared library loads, exports every symbol include/b200q.h declares, and its host-only
entry points (shard rule, argument validation, error strings) behave -- no compute call is made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from blazr_b200 import ops

HEADER = os.path.join(os.path.dirname(__file__), "..", "include", "b200q.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = ops.lib()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in b200q.h but not exported"
    assert sorted(ops.EXPORTS) == syms


def test_version_and_error_string():
    L = ops.lib()
    assert L.b200q_version() == 100
    assert isinstance(L.b200q_last_error(), bytes)


def test_shard_range_matches_reference_rule_and_oracle():
    # golden values of reference src/engine/tensor_parallel.rs:169-185
    assert ops.shard_range(32, 0, 4) == (0, 8) and ops.shard_range(32, 3, 4) == (24, 32)
    assert ops.shard_range(10, 0, 3) == (0, 4) and ops.shard_range(10, 2, 3) == (7, 10)
    for total in (0, 5, 64, 1002):
        for world in (1, 2, 4, 8):
            for r in range(world):
                assert ops.shard_range(total, r, world) == oracle.shard_range(total, r, world)
    # block-granular: 70B o_proj K=8192 rows split over 8 ranks at 256 granularity -> 4 super-blocks each
    assert ops.shard_range(8192, 3, 8, granule=256) == (3072, 4096)
    # 28672 = 112 super-blocks over 8 ranks -> 14 each
    assert ops.shard_range(28672, 7, 8, granule=256) == (25088, 28672)
    with pytest.raises(ops.B200QError):
        ops.shard_range(100, 0, 2, granule=256)
    with pytest.raises(ops.B200QError):
        ops.shard_range(10, 3, 3)


def test_invalid_arguments_return_codes_not_crashes():
    L = ops.lib()
    h = C.c_void_p()
    blocks = np.zeros(144, dtype=np.uint8)
    # unsupported ggml type -> UNSUPPORTED (no CPU fallback), message mentions it
    rc = L.b200q_weight_from_ggml(C.c_int32(9), C.c_void_p(blocks.ctypes.data), C.c_int32(0), C.c_int64(1), C.c_int64(256),
                                  C.c_int32(0), None, C.byref(h))
    assert rc == -2 and b"no CPU fallback" in L.b200q_last_error()
    # K not a multiple of the block
    rc = L.b200q_weight_from_ggml(C.c_int32(12), C.c_void_p(blocks.ctypes.data), C.c_int32(0), C.c_int64(1), C.c_int64(100),
                                  C.c_int32(0), None, C.byref(h))
    assert rc == -1
    # null out pointer
    rc = L.b200q_weight_from_ggml(C.c_int32(12), C.c_void_p(blocks.ctypes.data), C.c_int32(0), C.c_int64(1), C.c_int64(256),
                                  C.c_int32(0), None, None)
    assert rc == -1
    assert L.b200q_weight_free(None) == 0
    assert L.b200q_workspace_bytes(None, 1) == 0
    assert L.b200q_matmul(None, None, 0, C.c_int64(1), C.c_int64(1), None, 0, C.c_int64(1), None, C.c_size_t(0), None) == -1
    assert L.b200q_act_bytes(C.c_int64(4096), C.c_int64(1)) == 16 * 320
    assert L.b200q_act_bytes(C.c_int64(4100), C.c_int64(2)) == 17 * 2 * 320


def test_no_gpu_means_error_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ops.B200QError):
        ops.B200Client(0)
    L = ops.lib()
    assert L.b200q_device_count() == 0
    h = C.c_void_p()
    blocks = np.zeros(144 * 2, dtype=np.uint8)
    rc = L.b200q_weight_from_ggml(C.c_int32(12), C.c_void_p(blocks.ctypes.data), C.c_int32(0), C.c_int64(2), C.c_int64(256),
                                  C.c_int32(0), None, C.byref(h))
    assert rc == -3 and not h.value  # CUDA error surfaced, nothing silently computed on the host


def _sk_begin(g, C_, G):
    return g * C_ // G


def _sk_owner(c, C_, G):
    return ((c + 1) * G - 1) // C_


def test_stream_k_partition_formulae():
    """matvec.cu sk_begin / sk_owner: ranges tile [0,C) exactly and owner() inverts begin()."""
    rng = np.random.default_rng(0)
    cases = [(1, 1), (7, 3), (512, 148), (448 * 16, 296), (32, 32), (1002 * 16, 296), (5, 4)]
    cases += [(int(c), int(g)) for c, g in zip(rng.integers(1, 5000, 40), rng.integers(1, 300, 40))]
    for C_, G in cases:
        G = min(G, C_)
        begins = [_sk_begin(g, C_, G) for g in range(G + 1)]
        assert begins[0] == 0 and begins[-1] == C_
        assert all(b1 > b0 for b0, b1 in zip(begins, begins[1:]))  # no empty CTA
        sizes = [b1 - b0 for b0, b1 in zip(begins, begins[1:])]
        assert max(sizes) - min(sizes) <= 1
        for g in range(G):
            for c in {begins[g], begins[g + 1] - 1}:
                assert _sk_owner(c, C_, G) == g


def _sk_plan(c0, c1, KC):
    """python mirror of matvec.cu sk_plan()"""
    t0, t1 = c0 // KC, (c1 - 1) // KC
    kc0 = c0 - t0 * KC
    if kc0 != 0 or c1 < (t0 + 1) * KC:
        head_end = min((t0 + 1) * KC, c1)
    else:
        head_end = c0
    nH = head_end - c0
    tail_begin = c1
    if c1 > head_end and c1 != (t1 + 1) * KC:
        tail_begin = max(t1 * KC, head_end)
    nT = c1 - tail_begin
    nF = tail_begin - head_end
    return dict(nH=nH, nT=nT, nF=nF, kcH=kc0, tH=t0, tT=t1, tF=head_end // KC, head_end=head_end, tail_begin=tail_begin)


def test_stream_k_segment_plan():
    """head / tail / full segmentation: covers the range once, full segment is whole tiles, and the slot a
    contributor writes (head=0, tail=1) is the slot the reducer reads (first contributor: 0 iff it starts on
    the tile boundary, every later contributor: 0)."""
    rng = np.random.default_rng(1)
    cases = [(16, 112, 148), (16, 32, 148), (16, 224, 148), (56, 32, 148), (1, 100, 148), (112, 1, 148), (8, 4, 148), (3, 7, 5)]
    cases += [(int(k), int(t), int(g)) for k, t, g in zip(rng.integers(1, 120, 30), rng.integers(1, 300, 30), rng.integers(1, 300, 30))]
    for KC, T, G in cases:
        C_ = KC * T
        G = min(G, C_)
        contrib = {}
        for g in range(G):
            c0, c1 = _sk_begin(g, C_, G), _sk_begin(g + 1, C_, G)
            sp = _sk_plan(c0, c1, KC)
            assert sp["nH"] >= 0 and sp["nT"] >= 0 and sp["nF"] >= 0
            assert sp["nH"] + sp["nT"] + sp["nF"] == c1 - c0
            assert sp["nF"] % KC == 0 and (sp["nF"] == 0 or sp["head_end"] % KC == 0)
            if sp["nH"]:
                assert sp["nH"] < KC and (c0 + sp["nH"] - 1) // KC == sp["tH"]
                contrib.setdefault(sp["tH"], []).append((g, 0))
            if sp["nT"]:
                assert sp["nT"] < KC and sp["tail_begin"] % KC == 0 and sp["tail_begin"] // KC == sp["tT"]
                contrib.setdefault(sp["tT"], []).append((g, 1))
        for t, lst in contrib.items():
            gf, gl = _sk_owner(t * KC, C_, G), _sk_owner((t + 1) * KC - 1, C_, G)
            assert [g for g, _ in lst] == list(range(gf, gl + 1)), (KC, T, G, t, lst)
            slot_gf = 0 if _sk_begin(gf, C_, G) == t * KC else 1
            assert lst[0][1] == slot_gf and all(sl == 0 for _, sl in lst[1:])
            assert len(lst) >= 2  # a tile with a single contributor is never "partial"
