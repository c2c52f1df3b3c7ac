"""CPU check of the DEVICE tile-layout code: blazr_b200/csrc/formats.cuh (repack_row / load_unit of every format)
is compiled for the host (hostcheck/host_shim.h) and must reproduce the oracle's decomposition -- integer weights
q = v - off and the per-sub-block affine map (a, b) -- bit for bit.  Every kernel (dequantize, integer partials,
matvec, grouped matvec, tcgen05 GEMM) consumes a format only through that micro-interface, so this pins the
format-specific part of the GPU path without a GPU, including ragged shapes (N % 128 != 0, K % 256 != 0)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from blazr_b200 import synth

HC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "blazr_b200", "csrc", "hostcheck")
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ggml_dequant.npz"))


@pytest.fixture(scope="module")
def hostlib():
    so = os.path.join(HC, "libhostformats.so")
    srcs = [os.path.join(HC, "host_formats.cpp"), os.path.join(HC, "host_shim.h"), os.path.join(HC, "..", "formats.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas", "-o", so,
                        os.path.join(HC, "host_formats.cpp")], check=True, capture_output=True)
    return C.CDLL(so)


NATIVE = ["Q4_K", "Q6_K", "Q8_0", "Q5_K", "Q4_1", "Q5_1", "Q2_K", "Q3_K", "IQ4_XS", "TQ2_0"]
ADAPTED = ["Q4_0", "Q5_0", "IQ4_NL", "TQ1_0", "IQ2_XXS", "IQ2_XS", "IQ3_XXS", "IQ2_S", "IQ3_S", "IQ1_S", "IQ1_M"]


def _device_decompose(lib, name, blocks, N, K):
    t = synth.GGML[name]
    sub = oracle.sub_width(t)
    q = np.zeros((N, K), dtype=np.int8)
    a = np.zeros((N, K // sub), dtype=np.float32)
    b = np.zeros((N, K // sub), dtype=np.float32)
    fn = lib.hostfmt_decompose_adapted if name in ADAPTED else lib.hostfmt_decompose
    rc = fn(C.c_int(t), blocks.ctypes.data_as(C.c_void_p), C.c_int64(N), C.c_int64(K), q.ctypes.data_as(C.c_void_p),
            a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return q, a, b


@pytest.mark.parametrize("name", NATIVE + ADAPTED)
@pytest.mark.parametrize("shape", [(128, 256), (200, 768), (5, 2048), (130, 1024)])
def test_device_layout_matches_oracle_decomposition(hostlib, name, shape):
    N, K = shape
    t = synth.GGML[name]
    blocks = np.ascontiguousarray(synth.random_ggml(t, N, K, seed=N + K))
    q, a, b = _device_decompose(hostlib, name, blocks, N, K)
    qi, ra, rb, sub = oracle.decompose_ggml(t, blocks, N, K)
    assert np.array_equal(q, qi)
    assert np.array_equal(a.view(np.uint32), ra.view(np.uint32))
    assert np.array_equal(b.view(np.uint32), rb.view(np.uint32)) or np.array_equal(b, rb)  # +0.0 / -0.0 of an absent min


@pytest.mark.parametrize("name", ["Q8_0", "Q4_0", "Q5_0", "Q4_1", "Q5_1", "IQ4_NL"])
def test_ragged_k_for_32_element_blocks(hostlib, name):
    """K = 1408 (DeepSeek-V2-Lite expert width, 5.5 chunks): the last chunk is half valid"""
    N, K = 70, 1408
    t = synth.GGML[name]
    blocks = np.ascontiguousarray(synth.random_ggml(t, N, K, seed=9))
    q, a, b = _device_decompose(hostlib, name, blocks, N, K)
    qi, ra, rb, sub = oracle.decompose_ggml(t, blocks, N, K)
    assert np.array_equal(q, qi) and np.array_equal(a, ra) and np.array_equal(b, rb)


@pytest.mark.parametrize("name", NATIVE + ADAPTED)
def test_device_layout_dequantizes_the_golden_fixture(hostlib, name):
    """w = a (v - off) - b with separate multiply and subtract (contract #1) reproduces gguf.quants on the committed
    golden blocks (incl. all-0x00 / all-0xFF payloads) through the device layout code"""
    blk, deq = GOLD[f"{name}_blocks"], GOLD[f"{name}_deq"]
    N, K = deq.shape
    t = synth.GGML[name]
    sub = oracle.sub_width(t)
    q, a, b = _device_decompose(hostlib, name, np.ascontiguousarray(blk), N, K)
    w = (np.repeat(a, sub, axis=1) * q.astype(np.float32)).astype(np.float32) - np.repeat(b, sub, axis=1)
    assert np.array_equal(w.astype(np.float32).view(np.uint32), deq.view(np.uint32))
