"""GPU parity of the formats added after the GPU budget of round 1 was spent: Q5_K, Q4_1, Q5_1, Q2_K, Q3_K, IQ4_XS, TQ2_0, TQ1_0, IQ2_XXS, IQ2_XS, IQ3_XXS, IQ2_S, IQ3_S, IQ1_S, IQ1_M.

Their format-specific code (repack_row / load_unit in formats.cuh) is verified bit for bit on the CPU
(tests/test_host_formats.py); every kernel they run through is format-generic and green on hardware for the other nine
formats.  These tests are the same contracts (#1 dequantized weights, #2 integer partials, #3 matvec, GEMM tolerance)
but have NOT run on a B200 yet, so they are marked xfail(strict=False): a pass shows as XPASS, a failure does not stop
the suite (-x) in front of verified tests.  The file name sorts last for the same reason.  Remove the mark once seen green.
"""
import pytest

import test_gpu_gemm as tg
import test_gpu_quant as tq

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="added after the round-1 GPU budget was spent: CPU-verified layouts, not yet run on hardware")]

NEW = ["Q5_K", "Q4_1", "Q5_1", "Q2_K", "Q3_K", "IQ4_XS", "TQ2_0", "TQ1_0", "IQ2_XXS", "IQ2_XS", "IQ3_XXS", "IQ2_S", "IQ3_S", "IQ1_S", "IQ1_M"]


@pytest.mark.parametrize("name", NEW)
def test_dequant_golden_fixture_bit_exact(client, name):
    tq.test_dequant_golden_fixture_bit_exact(client, name)


@pytest.mark.parametrize("fmt", NEW)
@pytest.mark.parametrize("shape", [(128, 256), (200, 768), (8, 2048)])
def test_dequant_bit_exact_vs_oracle(client, fmt, shape):
    tq.test_dequant_bit_exact_vs_oracle(client, fmt, shape)


@pytest.mark.parametrize("fmt", NEW)
def test_int_partials_bit_exact(client, fmt):
    tq.test_int_partials_bit_exact(client, fmt)


@pytest.mark.parametrize("fmt", NEW)
@pytest.mark.parametrize("shape", tq.SHAPES)
def test_matvec_m1_vs_oracle(client, fmt, shape):
    tq.test_matvec_m1_vs_oracle(client, fmt, shape)


@pytest.mark.parametrize("fmt", NEW)
@pytest.mark.parametrize("M", [2, 4])
def test_matvec_small_batch(client, fmt, M):
    tq.test_matvec_small_batch(client, fmt, M)


@pytest.mark.parametrize("fmt", NEW)
@pytest.mark.parametrize("NKM", [(128, 256, 32), (200, 768, 100), (384, 4096, 256)])
def test_gemm_vs_oracle(client, fmt, NKM):
    tg.test_gemm_vs_oracle(client, fmt, NKM)


# ---- EXPERIMENTAL persistent op-list kernel (csrc/dstep_impl.cuh): same status -- written after the GPU budget was spent ----
# A persistent kernel with grid barriers can hang if it is wrong, and a hung kernel cannot be interrupted by pytest: the test
# only runs when asked for (B200Q_TEST_EXPERIMENTAL=1, under `timeout`), never in the default -m gpu suite.
@pytest.mark.skipif(__import__("os").environ.get("B200Q_TEST_EXPERIMENTAL", "0") == "0", reason="experimental persistent kernel: opt-in (B200Q_TEST_EXPERIMENTAL=1)")
@pytest.mark.parametrize("mode", ["1", "2"])
@pytest.mark.parametrize("scheme", ["Q6_K", "Q4_K_M", "Q8_0", "AWQ"])
def test_dstep_programs_reproduce_the_launch_per_op_step(client, scheme, mode, monkeypatch):
    """B200Q_DSTEP=1: 3 launches per layer (program, attention, program); B200Q_DSTEP=2: the whole step in ONE launch.
    Both must give the same logits, bit for bit, as the 8 separate launches per layer (identical arithmetic), over several
    dependent steps; then under CUDA-graph replay"""
    import numpy as np
    import torch
    from blazr_b200 import decode
    cfg = decode.PRESETS["tiny"]
    hm = decode.build_host_model(cfg, scheme, seed=31)
    ref = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    monkeypatch.setenv("B200Q_DSTEP", mode)
    exp = decode.Decoder(client, cfg, scheme, batch=1, max_ctx=64, host=hm)
    assert (exp.programs is not None) if mode == "1" else (exp.step_program is not None)
    for d in (ref, exp):
        d.reset([3])
    for _ in range(6):
        ref.step()
        exp.step()
        torch.cuda.synchronize()
        a, b = ref.logits.cpu().numpy(), exp.logits.cpu().numpy()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    prompt = np.asarray([[3, 1, 4, 1, 5]])
    assert np.array_equal(ref.generate(prompt, 16, use_graph=True), exp.generate(prompt, 16, use_graph=True))
