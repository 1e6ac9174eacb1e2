"""oracle/crt_gemm.py — the CPU restatement of the integer-residue (CRT) FP64 contraction of compute mode 'i8crt' — against
exact integer arithmetic and against a plain FP64 matmul."""
import math

import numpy as np
import pytest

from oracle import crt_gemm as G


def test_moduli_are_pairwise_coprime_and_fit_int8():
    for i, p in enumerate(G.MODULI):
        assert 2 <= p <= 256
        for q in G.MODULI[:i]:
            assert math.gcd(p, q) == 1


@pytest.mark.parametrize('T,k,bits', [(15, 1024, 53), (15, 2048, 52), (16, 16384, 53), (16, 65536, 53), (9, 1024, 30)])
def test_bit_budget(T, k, bits):
    assert G.max_bits(T, k) == bits


@pytest.mark.parametrize('T', [16, 15, 12, 9])
def test_pipeline_equals_the_exact_product_of_the_truncated_operands(T):
    rng = np.random.default_rng(T)
    A = rng.standard_normal((24, 300)) * np.exp(4 * rng.standard_normal((24, 1)))
    B = rng.standard_normal((17, 300)) * np.exp(4 * rng.standard_normal((17, 1)))
    A[3] = 0.0                                    # an all-zero row
    A[5, 7] = 2.0 ** 10                           # a row whose maximum is an exact power of two
    b = G.max_bits(T, 300)
    C = G.crt_matmul(A, B, T)
    ex = G.exact_matmul_of_truncated(A, B, b)
    assert np.array_equal(C, ex)                  # residues + CRT words + scaling reproduce the exact integer product
    ref = A @ B.T
    err = np.linalg.norm(C - ref) / np.linalg.norm(ref)
    assert err < {16: 5e-15, 15: 5e-15, 12: 1e-11, 9: 2e-8}[T]


def test_cancellation_heavy_products_keep_fp64_accuracy():
    """Rows of an inverse Cholesky factor against kernel columns: terms of size 1e3 summing to O(1)."""
    rng = np.random.default_rng(1)
    Z = rng.standard_normal((200, 3))
    Kzz = 1.5 * np.exp(-0.5 * ((Z[:, None, :] - Z[None, :, :]) ** 2).sum(-1) / 4.0) + 1e-6 * np.eye(200)
    Linv = np.linalg.inv(np.linalg.cholesky(Kzz))
    X = rng.standard_normal((64, 3))
    K = 1.5 * np.exp(-0.5 * ((X[:, None, :] - Z[None, :, :]) ** 2).sum(-1) / 4.0)
    C = G.crt_matmul(K, Linv, 15)
    import mpmath
    mpmath.mp.dps = 40
    exact = np.array((mpmath.matrix(K.tolist()) * mpmath.matrix(Linv.T.tolist())).tolist(), dtype=float)
    e_crt = np.linalg.norm(C - exact) / np.linalg.norm(exact)
    e_f64 = np.linalg.norm(K @ Linv.T - exact) / np.linalg.norm(exact)
    assert e_crt < 4 * e_f64 + 1e-16, (e_crt, e_f64)
