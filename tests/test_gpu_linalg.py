"""GPU svd3 / polar / determinant against the oracle and the reference's gtest cases
(reference tests/test_linalg.cu:25-91)."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu

M1 = np.array([[1, 2, 1], [1, 3, 1], [1, 8, 1]], np.float32)             # test_linalg.cu:27
M2 = np.array([[0, 1, 0], [-1, 2, -1], [-1, 0.001, -1]], np.float32)     # test_linalg.cu:33


def _matrices(n, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, 9)).astype(np.float32)
    q = n // 4
    A[:q] = np.eye(3, dtype=np.float32).reshape(9) + 0.05 * A[:q]
    A[q:2 * q] *= np.float32(10) ** rng.uniform(-6, 6, (q, 1)).astype(np.float32)
    A[-100:, 2] = A[-100:, 0]; A[-100:, 5] = A[-100:, 3]; A[-100:, 8] = A[-100:, 6]  # rank deficient
    A[-200:-100] = 0
    A[-300:-200] = np.diag([1, 1, -0.9]).astype(np.float32).reshape(9)  # inverted
    A[0] = M1.reshape(9)
    A[1] = M2.reshape(9)
    return A


def test_svd3_exact_is_bit_identical_to_reference_arithmetic():
    import mpm_b200

    A = _matrices(1_000_000)
    U, S, V = mpm_b200.svd3_batch(A, mpm_b200.SVD_EXACT)
    Uo, So, Vo = ol.svd3(A)
    for g, o in ((U, Uo), (S, So), (V, Vo)):
        same = (g.view(np.uint32) == o.view(np.uint32)) | (np.isnan(g) & np.isnan(o))
        assert same.all(), f"{(~same).sum()} mismatching words"


def test_svd3_fast_deviation_from_exact():
    import mpm_b200

    A = _matrices(200_000, seed=3)[: 100_000]  # near-identity + scaled blocks, the physical regime
    A = A[:50_000]
    U, S, V = mpm_b200.svd3_batch(A, mpm_b200.SVD_FAST)
    Uo, So, Vo = ol.svd3(A)
    rec = np.einsum("nij,nj,nkj->nik", U.astype(np.float64), S.astype(np.float64), V.astype(np.float64))
    err = np.abs(rec - A.reshape(-1, 3, 3)).sum((1, 2))
    assert err.max() < 1e-5  # same L1 bound the reference's gtest uses
    assert np.abs(S - So).max() < 5e-6
    R, Ro = np.einsum("nij,nkj->nik", U, V), np.einsum("nij,nkj->nik", Uo, Vo)
    assert np.abs(R - Ro).max() < 2e-5


@pytest.mark.parametrize("M", [M1, M2])
@pytest.mark.parametrize("mode", [0, 1])
def test_polar_gtest_cases(M, mode):
    """TestDevicePolar.Basic / TestPolar.*: |RS - M|_1 < 1e-5, R unitary, S symmetric."""
    import mpm_b200

    U, S, V = mpm_b200.svd3_batch(M, mode)
    U, S, V = U[0].astype(np.float32), S[0], V[0].astype(np.float32)
    R = mpm_b200.polar_batch(M, mode)[0]
    Sym = (V * S) @ V.T
    assert np.abs(R @ Sym - M).sum() < 1e-5
    assert np.abs(R @ R.T - np.eye(3)).sum() < 1e-5
    assert np.abs(Sym - Sym.T).sum() < 1e-5


def test_determinant_identity():
    """TestDeterminant.Identity: det(I) == 1.0 exactly."""
    import mpm_b200

    d = mpm_b200.determinant_batch(np.stack([np.eye(3, dtype=np.float32), M1]))
    assert d[0] == 1.0
    assert d[1] == ol.determinant(M1)
