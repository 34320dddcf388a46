"""CPU tests: the oracle against the reference's golden vectors, the committed fixtures made from
the reference's own code (tests/make_golden.py), the live oracle/_ref build where present, and
oracle-free invariants."""
import os

import numpy as np
import pytest

import oracle_lib as ol
import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M1 = np.array([[1, 2, 1], [1, 3, 1], [1, 8, 1]], np.float32)          # reference tests/test_linalg.cu:27
M2 = np.array([[0, 1, 0], [-1, 2, -1], [-1, 0.001, -1]], np.float32)  # :33


def _l1(a):
    return float(np.abs(a).sum())


@pytest.mark.parametrize("M", [M1, M2])
def test_polar_gtest_bounds(M):
    """TestPolar.Basic/Harder/UnitaryHermitian + TestDevicePolar.Basic (test_linalg.cu:10-75)."""
    R, S = ol.polar(M)
    R, S = R[0], S[0]
    assert _l1(R @ S - M) < 1e-5
    assert _l1(R @ R.T - np.eye(3, dtype=np.float32)) < 1e-5
    assert _l1(S - S.T) < 1e-5


def test_determinant_identity():
    """TestDeterminant.Identity (test_linalg.cu:77-91)."""
    assert ol.determinant(np.eye(3, dtype=np.float32)) == 1.0


def test_svd3_known_answers():
    U, S, V = ol.svd3(np.stack([M1, M2, np.eye(3, dtype=np.float32), np.diag([1, 1, -0.9]).astype(np.float32)]))
    assert np.allclose(S[0], [9.026522, 1.233645, 0.0], atol=2e-6)
    assert np.allclose(S[1], [2.715451, 1.275275, 0.0], atol=2e-6)
    assert np.array_equal(S[2], [1, 1, 1]) and np.array_equal(U[2], np.eye(3)) and np.array_equal(V[2], np.eye(3))
    assert S[3][2] == np.float32(-0.9)  # negative-sigma convention: U, V stay rotations


def test_svd3_golden_bit_exact():
    g = np.load(os.path.join(GOLD, "svd3_golden.npz"))
    U, S, V = ol.svd3(g["A"])
    for mine, ref in ((U, g["U"]), (S, g["S"]), (V, g["V"])):
        same = (mine.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(mine) & np.isnan(ref))
        assert same.all()


def test_svd3_live_reference_bit_exact():
    if ol.ref_lib("libref_svd3") is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(11)
    A = rng.standard_normal((300_000, 9)).astype(np.float32)
    A[:100_000] = np.eye(3, dtype=np.float32).reshape(9) + 0.1 * A[:100_000]
    a, b = ol.svd3(A), ol.svd3(A, which="ref")
    for x, y in zip(a, b):
        assert (x.view(np.uint32) == y.view(np.uint32)).all()


@pytest.mark.parametrize("name,kind", [("snow", ol.SNOW), ("fc", ol.FIXED_COROTATED), ("jelly", ol.JELLY)])
def test_substep_golden_bit_exact(name, kind):
    """Oracle == the reference's own plugin headers + kernel bodies (host build), bit for bit,
    stage by stage and over a 21-step horizon."""
    g = np.load(os.path.join(GOLD, f"substep_{name}_golden.npz"))
    N, dt, mats = int(g["N"]), float(g["dt"]), g["mats"]
    p = g["p0"].copy().view(ol.PARTICLE_DTYPE).reshape(-1)
    ol.set_threads(1)  # serial P2G = the fixture's summation order
    grid = ol.p2g(p, mats, dt, N, kind)
    assert grid.tobytes() == g["grid_after_p2g"].tobytes()
    ol.grid_update(grid, dt, N)
    assert grid.tobytes() == g["grid_after_update"].tobytes()
    ol.g2p(grid, p, mats, dt, N, kind)
    assert p.tobytes() == g["p1"].tobytes()
    ol.advance(p, mats, dt, N, kind, 20)
    assert p.tobytes() == g["p21"].tobytes()
    ol.set_threads(ol.max_threads())


@pytest.mark.parametrize("kind", [ol.SNOW, ol.FIXED_COROTATED, ol.JELLY])
def test_substep_live_reference(kind):
    ref = ol.Ref(kind)
    if not ref.available:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    N, dt = 24, 1e-4
    p, mats = scenes.two_spheres(N, density=100000.0, seed=3, kind=kind)
    assert np.array_equal(ref.make_material(1 / 100000.0), ol.make_material(1 / 100000.0))
    assert ref.params(dt, 60) == ol.params(dt, 60)
    a, b = p.copy(), p.copy()
    nthreads = ol.max_threads()
    ol.set_threads(1)
    ol.advance(a, mats, dt, N, kind, 10)
    ref.advance(b, mats, dt, N, 10)
    ol.set_threads(nthreads)
    assert a.tobytes() == b.tobytes()


def test_params_non_power_of_two():
    dx, dx_inv = ol.params(1e-4, 60)  # dx_inv is not exactly N
    assert dx == np.float32(1.0 / 60) and abs(dx_inv - 59.9999962) < 1e-5


def test_wall_planes():
    """Which planes are sticky (App. A table): N=16 has no high wall, N=64: i<=3 and i>=61."""
    for N, lo, hi in ((16, 0, None), (32, 1, 31), (64, 3, 61)):
        g = ol.new_grid(N)
        g[...] = 1.0
        ol.grid_update(g, 0.0, N)
        sticky = np.where(g[:, N // 2, N // 2, 0] == 0)[0]
        lows = sticky[sticky < N // 2]
        highs = sticky[sticky >= N // 2]
        assert lows.max() == lo
        assert (highs.min() == hi) if hi is not None else len(highs) == 0


def test_weights_partition_of_unity_and_moments():
    rng = np.random.default_rng(0)
    N = 60
    dx, dx_inv = ol.params(1e-4, N)
    for x in rng.uniform(0.05, 0.95, (200, 3)).astype(np.float32):
        base, w = ol.weights(x, dx_inv)
        assert np.allclose(w.sum(1), 1.0, atol=1e-6)
        d = (base[:, None] + np.arange(3)[None, :]) * dx - x[:, None]
        assert np.allclose((w * d).sum(1), 0.0, atol=1e-7)                  # sum w d = 0
        assert np.allclose((w * d * d).sum(1), dx * dx / 4, rtol=2e-4)      # D = dx^2/4 I  <=> D_inv_const


def test_p2g_conserves_mass_and_momentum():
    N, dt = 32, 1e-4
    p, mats = scenes.two_spheres(N, kind=ol.SNOW)
    p["F"][:, :] = 0
    p["F"][:, [0, 4, 8]] = 1
    p["Jp"] = 1  # no stress: grid momentum = sum m v
    g = ol.p2g(p, mats, dt, N, ol.SNOW)
    m = float(mats[1])
    assert abs(g[..., 3].sum(dtype=np.float64) - m * len(p)) < 1e-6 * m * len(p)
    mom = (p["v"].astype(np.float64) * m).sum(0)
    assert np.allclose(g[..., :3].reshape(-1, 3).sum(0, dtype=np.float64), mom, rtol=1e-4, atol=1e-7)


def test_sort_keys_and_perm():
    N = 16
    x = np.array([[0.5, 0.5, 0.5], [0.01, 0.99, 0.5], [-0.2, 0.5, 1.3], [0.5, 0.5, 0.5]], np.float32)
    p = ol.new_particles(x)
    k = ol.cell_keys(p, 1e-4, N)
    assert k[0] == k[3] == (7 * 16 + 7) * 16 + 7
    assert k[2] == (0 * 16 + 7) * 16 + 15  # clamped base node
    perm = ol.sort_perm(k)
    assert list(perm) == sorted(range(4), key=lambda i: (k[i], i))


def test_openmp_matches_serial_within_rounding():
    N, dt = 32, 1e-4
    p, mats = scenes.two_spheres(N, kind=ol.SNOW)
    nthreads = ol.max_threads()
    ol.set_threads(1)
    a = ol.p2g(p, mats, dt, N, ol.SNOW)
    ol.set_threads(nthreads)
    b = ol.p2g(p, mats, dt, N, ol.SNOW)
    scale = np.abs(a).max((0, 1, 2))
    assert (np.abs(a - b).max((0, 1, 2)) <= 1e-5 * scale).all()
