"""CPU check of the matvec kernel's stream-K plan (blazr_b200/csrc/streamk_plan.cuh, compiled for the host): for every
grid size and (tiles, k-chunks) combination -- including every projection shape of the BASELINE configs at TP 1/2/4/8 and
the grouped MoE launches -- each chunk is streamed exactly once, the consumer walk flushes the tile it accumulated, and
the fix-up's contributor range / count / partial-slot bookkeeping matches the CTAs that really contribute, and the wide
reducer's premise holds (contributor gf + 1 of a tile split three or more ways owns chunks of that tile only)."""
import ctypes as C
import os
import subprocess

import pytest

HC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "blazr_b200", "csrc", "hostcheck")


@pytest.fixture(scope="module")
def sk():
    so = os.path.join(HC, "libhoststreamk.so")
    srcs = [os.path.join(HC, "host_streamk.cpp"), os.path.join(HC, "..", "streamk_plan.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HC, "host_streamk.cpp")], check=True, capture_output=True)
    return C.CDLL(so)


def _check(sk, T, KC, G):
    ns, mc = C.c_int64(), C.c_int64()
    rc = sk.streamk_check(C.c_int64(T), C.c_int64(KC), C.c_int64(G), C.byref(ns), C.byref(mc))
    assert rc == 0, (T, KC, G, rc)
    return ns.value, mc.value


def test_exhaustive_small(sk):
    for T in range(1, 25):
        for KC in range(1, 20):
            for G in (1, 2, 3, 5, 7, 8, 13, 16, 31, 64, 148):
                _check(sk, T, KC, G)


def test_model_shapes_at_every_tp_degree(sk):
    """(N, K) of q|k|v, o, gate|up, down, lm_head of the BASELINE models, sharded as blazr_b200/tp.py does"""
    models = [(2048, 32, 8, 64, 8192, 128256), (4096, 32, 8, 128, 14336, 32000), (4096, 32, 8, 128, 14336, 128256),
              (8192, 64, 8, 128, 28672, 128256)]
    for hidden, nh, nkv, hd, ffn, vocab in models:
        for world in (1, 2, 4, 8):
            shapes = [((nh + 2 * nkv) * hd // world, hidden), (hidden, nh * hd // world), (2 * ffn // world, hidden),
                      (hidden, ffn // world), (-(-vocab // world), hidden)]
            for N, K in shapes:
                T, KC = -(-N // 128), -(-K // 256)
                ns, mc = _check(sk, T, KC, 148)
                assert mc <= KC + 1


def test_grouped_moe_launches(sk):
    """tiles numbered slot-major: n_slots x tiles-per-expert; DeepSeek-V2-Lite expert shapes, 1..64 slots"""
    for n_slots in (1, 2, 6, 8, 48, 64):
        _check(sk, n_slots * 22, 8, 148)   # gate|up 2816 x 2048
        _check(sk, n_slots * 16, 6, 148)   # down 2048 x 1408 (K padded to 1536)


def test_random_large(sk):
    import random
    rnd = random.Random(7)
    for _ in range(300):
        _check(sk, rnd.randint(1, 1100), rnd.randint(1, 120), rnd.choice([132, 148, 160, 37]))
