"""GPU parity of the formats added after the GPU budget of round 1 was spent: Q5_K, Q4_1, Q5_1, Q2_K, Q3_K, IQ4_XS.

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

NEW = ["Q5_K", "Q4_1", "Q5_1", "Q2_K", "Q3_K", "IQ4_XS"]


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
