"""GPU parity tests (call through the C ABI): repack + dequantize, activation quantiser, integer partials
and the dp4a stream-K matvec against the CPU oracle on identical synthetic packed weights.

Contracts (BASELINE.json north_star / SURVEY.md section 8c):
  #1 dequantized weights bit-exact          #2 integer dot partials bit-exact
  #3 Y within 1e-2 relative (max-abs / max-abs) of the f32 flavour A; we additionally hold the device
     result within 2e-4 of the oracle's int8-activation flavour B (same arithmetic, different f32 order).
"""
import os

import numpy as np
import pytest
import torch

import oracle
from blazr_b200 import ops, synth
from qcases import ALL, GGML_GPU, Case

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ggml_dequant.npz"))
TOL_A = 1e-2   # north_star: 1e-2 relative on outputs vs the reference CPU (f32) path
TOL_B = 1e-6   # same int8 arithmetic with exact f64 accumulation on both sides: bit-identical up to rare 1-ulp cases


def rel_err(y, ref):
    return float(np.abs(y.astype(np.float64) - ref.astype(np.float64)).max() / max(np.abs(ref).max(), 1e-30))


def assert_bit_exact(y, ref, what=""):
    """contract #3 (decode path): the f32 outputs equal the oracle's bit for bit; a double-rounding
    coincidence (f64 sum within 1e-16 of an f32 midpoint) may flip the last bit of < 1e-4 of the elements"""
    y, ref = np.asarray(y, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    neq = y.view(np.uint32) != ref.view(np.uint32)
    assert neq.mean() < 1e-4, (what, float(neq.mean()))
    assert rel_err(y, ref) < TOL_B, what


def test_native_library_is_loaded(client):
    assert os.path.exists(ops.LIB_PATH)
    assert ops.lib().b200q_device_count() >= 1
    maps = open("/proc/self/maps").read()
    assert "libb200q.so" in maps


# ---------------------------------------------------------------- contract #1: dequantize
@pytest.mark.parametrize("name", GGML_GPU)
def test_dequant_golden_fixture_bit_exact(client, name):
    """device repack + dequantize reproduces gguf.quants on the committed golden blocks"""
    t = synth.GGML[name]
    blk, deq = GOLD[f"{name}_blocks"], GOLD[f"{name}_deq"]
    N, K = deq.shape
    w = client.weight_from_ggml(t, blk, N, K)
    out = client.dequantize(w).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), deq.view(np.uint32))


@pytest.mark.parametrize("fmt", ALL)
@pytest.mark.parametrize("shape", [(128, 256), (200, 768), (8, 2048), (1408, 512)])
def test_dequant_bit_exact_vs_oracle(client, fmt, shape):
    N, K = shape
    c = Case(client, fmt, N, K, seed=N + K)
    out = client.dequantize(c.w).cpu().numpy()
    assert out.shape == (N, K)
    assert np.array_equal(out.view(np.uint32), c.deq.view(np.uint32))
    # half outputs are the f32 value rounded once
    out16 = client.dequantize(c.w, torch.float16).float().cpu().numpy()
    assert np.array_equal(out16, c.deq.astype(np.float16).astype(np.float32))


def test_q8_0_ragged_k_is_zero_padded(client):
    """K = 1408 (DeepSeek-V2-Lite expert down_proj) is 5.5 super-blocks: 32-block formats pad to 256"""
    N, K = 256, 1408
    c = Case(client, "Q8_0", N, K, seed=9)
    assert c.w.K_pad == 1536
    out = client.dequantize(c.w).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), c.deq.view(np.uint32))
    x = synth.random_act(1, K)
    y = client.quant_matmul(torch.from_numpy(x).cuda(), c.w).cpu().numpy()
    assert rel_err(y, c.oracle_b(x)) < TOL_B


def test_device_resident_source_blocks(client):
    """blazr uploads raw blocks to the device first (gguf.rs:33): accept a device pointer as the source"""
    t = synth.GGML["Q4_K"]
    blk = synth.random_ggml(t, 256, 512, seed=4)
    w = client.weight_from_ggml(t, torch.from_numpy(blk).cuda(), 256, 512)
    out = client.dequantize(w).cpu().numpy()
    assert np.array_equal(out.view(np.uint32), oracle.dequant_ggml(t, blk, 256, 512).view(np.uint32))


# ---------------------------------------------------------------- activation quantiser
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("MK", [(1, 256), (3, 4096), (4, 1408)])
def test_act_quant_bit_exact(client, dtype, MK):
    M, K = MK
    x = synth.random_act(M, K, seed=K)
    x[0, :32] = 0.0
    x[-1, 40] = 1e4  # an outlier block
    xt = torch.from_numpy(x).cuda().to(dtype)
    xq = client.quantize_act(xt)
    q, d, bs = (t.cpu().numpy() for t in client.act_unpack(xq, M, K))
    K_pad = (K + 255) // 256 * 256
    xo = np.zeros((M, K_pad), dtype=np.float32)
    xo[:, :K] = xt.float().cpu().numpy()
    qo, do, bso = oracle.quantize_act(xo)
    assert np.array_equal(q, qo)
    assert np.array_equal(d.view(np.uint32), do.view(np.uint32))
    assert np.array_equal(bs, bso)


# ---------------------------------------------------------------- contract #2: integer partials
@pytest.mark.parametrize("fmt", ALL)
def test_int_partials_bit_exact(client, fmt):
    N, K, M = 200, 1024, 2
    c = Case(client, fmt, N, K, seed=21)
    x = synth.random_act(M, K, seed=3)
    xt = torch.from_numpy(x).cuda()
    perm_t = torch.from_numpy(c.perm).cuda() if c.perm is not None else None
    xq = client.quantize_act(xt, perm_t)
    part = client.int_partials(c.w, xq, M).cpu().numpy()
    xq_o, _, _ = oracle.quantize_act(c.x_perm(x))
    ref = oracle.int_partials(c.qi, xq_o, c.sub)
    assert np.array_equal(part, ref)


# ---------------------------------------------------------------- contract #3: matvec outputs
SHAPES = [
    (128, 256),      # one chunk
    (512, 2048),     # Llama-3.2-1B k/v
    (2048, 2048),    # 1B q/o
    (1024, 4096),    # 7B/8B k/v
    (200, 768),      # ragged N
    (128, 8192),     # small N, long K: one tile split over many CTAs (stream-K fix-up path)
    (384, 14336),    # 7B down_proj slice: tiles split across CTAs
    (14336, 1024),   # many tiles, short K
]


@pytest.mark.parametrize("fmt", ALL)
@pytest.mark.parametrize("shape", SHAPES)
def test_matvec_m1_vs_oracle(client, fmt, shape):
    N, K = shape
    c = Case(client, fmt, N, K, seed=N * 7 + K)
    x = synth.random_act(1, K, seed=5)
    y = client.quant_matmul(torch.from_numpy(x).cuda(), c.w, path=ops.PATH_MATVEC).cpu().numpy()
    assert y.shape == (1, N)
    assert_bit_exact(y, c.oracle_b(x), (fmt, shape))
    assert rel_err(y, c.oracle_a(x)) < TOL_A


@pytest.mark.parametrize("fmt", ALL)
@pytest.mark.parametrize("M", [2, 3, 4])
def test_matvec_small_batch(client, fmt, M):
    N, K = 640, 2048
    c = Case(client, fmt, N, K, seed=M)
    x = synth.random_act(M, K, seed=6)
    y = client.quant_matmul(torch.from_numpy(x).cuda(), c.w, path=ops.PATH_MATVEC).cpu().numpy()
    assert_bit_exact(y, c.oracle_b(x), (fmt, M))
    assert rel_err(y, c.oracle_a(x)) < TOL_A


@pytest.mark.parametrize("fmt", ["AWQ", "GPTQ", "Q4_K"])
def test_matvec_half_activations_and_outputs(client, fmt):
    """AWQ/GPTQ models run with f16 activations (awq.rs:69-71, gptq.rs:66-68)"""
    N, K = 1024, 4096
    c = Case(client, fmt, N, K, seed=2)
    x16 = torch.from_numpy(synth.random_act(1, K)).cuda().half()
    y = client.quant_matmul(x16, c.w)
    assert y.dtype == torch.float16
    ref = c.oracle_b(x16.float().cpu().numpy())
    assert rel_err(y.float().cpu().numpy(), ref) < 2e-3  # one f16 rounding of the output
    ybf = client.quant_matmul(x16.bfloat16(), c.w, out_dtype=torch.float32).cpu().numpy()
    assert rel_err(ybf, c.oracle_b(x16.bfloat16().float().cpu().numpy())) < TOL_B


def test_gptq_bias_is_added(client):
    N, K = 256, 1024
    c = Case(client, "GPTQ", N, K, seed=8, bias=True)
    assert c.bias is not None and c.w.info.has_bias
    x = synth.random_act(2, K)
    y = client.quant_matmul(torch.from_numpy(x).cuda(), c.w).cpu().numpy()
    assert rel_err(y, c.oracle_b(x)) < TOL_B
    y0 = oracle.matmul_q8(c.qi, c.a, c.b, c.sub, x, None)
    assert np.abs((y - y0) - c.bias[None, :]).max() < 1e-3


def test_strided_input_and_output(client):
    N, K, M = 256, 512, 2
    c = Case(client, "Q6_K", N, K, seed=1)
    xbig = torch.from_numpy(synth.random_act(M, K + 64)).cuda()
    x = xbig[:, :K]                                   # ldx = K + 64
    ybig = torch.zeros((M, N + 32), device="cuda")
    client.quant_matmul(x, c.w, out=ybig[:, :N])      # ldy = N + 32
    ref = c.oracle_b(x.cpu().numpy())
    assert rel_err(ybig[:, :N].cpu().numpy(), ref) < TOL_B
    assert float(ybig[:, N:].abs().max()) == 0.0


def test_matvec_is_deterministic_and_workspace_self_cleaning(client):
    """stream-K partials are combined in CTA order: repeated calls on one workspace are bit-identical"""
    c = Case(client, "Q4_K", 128, 8192, seed=3)
    x = torch.from_numpy(synth.random_act(1, 8192)).cuda()
    ys = [client.quant_matmul(x, c.w).clone() for _ in range(5)]
    for y in ys[1:]:
        assert torch.equal(y, ys[0])


def test_matvec_linearity_at_full_size(client):
    """size-independent property at a BASELINE config shape (Mistral-7B gate_proj, Q6_K): the int8
    quantiser is scale-equivariant for powers of two, so W(2x) == 2 W(x) exactly."""
    N, K = 14336, 4096
    c = Case(client, "Q6_K", N, K, seed=77)
    x = torch.from_numpy(synth.random_act(1, K)).cuda()
    y1 = client.quant_matmul(x, c.w)
    y2 = client.quant_matmul(2.0 * x, c.w)
    assert torch.equal(y2, 2.0 * y1)
    assert rel_err(y1.cpu().numpy(), c.oracle_b(x.cpu().numpy())) < TOL_B


def test_cuda_graph_capture(client):
    """the op must be capturable (reference src/engine/cuda_graphs.rs:101-130): no alloc, no sync"""
    c = Case(client, "Q8_0", 512, 2048, seed=12)
    x = torch.from_numpy(synth.random_act(1, 2048)).cuda()
    y = torch.empty((1, 512), device="cuda")
    ws = c.w.workspace(1)
    client.quant_matmul(x, c.w, out=y, workspace=ws)
    torch.cuda.synchronize()
    ref = y.clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            client.quant_matmul(x, c.w, out=y, workspace=ws)
    y.zero_()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, ref)


def test_shared_activation_split_api(client):
    """q, k, v share one quantised activation (b200q_quantize_act + b200q_matmul_q8)"""
    K = 2048
    cq, ck = Case(client, "Q4_K", 2048, K, seed=1), Case(client, "Q6_K", 512, K, seed=2)
    x = synth.random_act(1, K)
    xq = client.quantize_act(torch.from_numpy(x).cuda())
    yq = client.matmul_q8(xq, 1, cq.w).cpu().numpy()
    yk = client.matmul_q8(xq, 1, ck.w).cpu().numpy()
    assert rel_err(yq, cq.oracle_b(x)) < TOL_B and rel_err(yk, ck.oracle_b(x)) < TOL_B


def test_errors_are_codes(client):
    c = Case(client, "Q4_K", 128, 256)
    x = torch.zeros((1, 256), device="cuda")
    with pytest.raises(ops.B200QError) as e:
        client.quant_matmul(x, c.w, workspace=torch.zeros(256, dtype=torch.uint8, device="cuda"))
    assert e.value.code == -4
    with pytest.raises(ops.B200QError) as e:
        client.weight_from_ggml(9, np.zeros(288, dtype=np.uint8), 1, 256)  # Q8_1 (activation-side format): no kernel, no fallback
    assert e.value.code == -2
    qw, sc, zr = synth.random_awq(128, 256)
    sc = sc * np.float32(1.0001)  # no longer f16-representable
    with pytest.raises(ops.B200QError):
        client.weight_from_decomposed(ops.DecomposedQuantTensor(qw, sc, zr, None, ops.DecomposedQuantMethod("awq", 128), (128, 256)))


def test_cpp_host_mirror_smoke(client):
    """the compiled-language host side (blazr_b200/host/b200q.hpp) drives the C ABI end to end"""
    import subprocess

    exe = os.path.join(os.path.dirname(ops.LIB_PATH), "..", "host", "host_smoke")
    if not os.path.exists(exe):
        pytest.skip("host_smoke not built")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_smoke ok" in r.stdout
