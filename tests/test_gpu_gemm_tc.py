"""The tcgen05 contraction kernel of the training step (csrc/ocf_gemm_tc.cuh) against float64 NumPy, through the C ABI:
all four operand arrangements (K-major / MN-major A and B), split-K over clusters of 1..8 CTAs reduced through
distributed shared memory, ragged sizes (zero-filled TMA boxes), plain tf32 and the fp32-grade 3-term split."""
import ctypes as C

import numpy as np
import pytest

from omnidirectional_collaborative_filtering_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(a_mn, b_mn, m, n, k, terms, split, seed):
    rs = np.random.RandomState(seed)
    A = rs.standard_normal((m, k)).astype(np.float32)          # A(m, k)
    B = rs.standard_normal((n, k)).astype(np.float32)          # B(n, k)
    a_host = np.ascontiguousarray(A.T if a_mn else A)
    b_host = np.ascontiguousarray(B.T if b_mn else B)
    out = np.full((n, m), np.nan, dtype=np.float32)
    _lib.check(_lib.lib().ocf_gemm_tc(_lib.ptr(a_host), int(a_mn), _lib.ptr(b_host), int(b_mn), m, n, k, terms, split, _lib.ptr(out)))
    want = B.astype(np.float64) @ A.astype(np.float64).T        # [n, m]
    scale = np.abs(B).astype(np.float64) @ np.abs(A).astype(np.float64).T
    return out, want, scale


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("m,n,k,split", [(128, 128, 32, 1), (256, 128, 256, 1), (128, 128, 1024, 8), (512, 128, 1024, 0),
                                         (384, 100, 224, 1), (128, 256, 512, 4), (1024, 512, 128, 0), (256, 64, 64, 2)])
def test_three_term_product_is_fp32_grade(a_mn, b_mn, m, n, k, split):
    out, want, scale = _run(a_mn, b_mn, m, n, k, 3, split, seed=m + n + k + 2 * a_mn + b_mn)
    assert np.isfinite(out).all()
    err = np.abs(out - want) / (scale + 1e-30)
    # fp32 accumulation of k products: a few ulp of the absolute-value product; tf32 alone would be ~5e-4
    assert err.max() < 4e-6, err.max()


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1)])
def test_plain_tf32_is_tf32_grade_and_worse(a_mn, b_mn):
    out1, want, scale = _run(a_mn, b_mn, 256, 128, 512, 1, 0, seed=3)
    out3, _, _ = _run(a_mn, b_mn, 256, 128, 512, 3, 0, seed=3)
    e1 = (np.abs(out1 - want) / scale).max()
    e3 = (np.abs(out3 - want) / scale).max()
    assert e1 < 2e-3 and e3 < 4e-6 and e1 > 20 * e3


def test_split_orders_are_reproducible():
    a, _, _ = _run(1, 0, 256, 128, 1024, 3, 8, seed=5)
    b, _, _ = _run(1, 0, 256, 128, 1024, 3, 8, seed=5)
    assert np.array_equal(a, b)


def test_bad_arguments_are_refused():
    x = np.zeros((8, 8), dtype=np.float32)
    lib = _lib.lib()
    assert lib.ocf_gemm_tc(_lib.ptr(x), 0, _lib.ptr(x), 0, 8, 8, 6, 3, 0, _lib.ptr(x)) != 0
    assert lib.ocf_gemm_tc(_lib.ptr(x), 0, _lib.ptr(x), 0, 8, 8, 8, 2, 0, _lib.ptr(x)) != 0
    assert lib.ocf_gemm_tc(_lib.ptr(x), 0, _lib.ptr(x), 0, 8, 8, 8, 3, 3, _lib.ptr(x)) != 0
