"""GPU substep (through the C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

import oracle_lib as ol
import scenes

pytestmark = pytest.mark.gpu

DT = 1e-4
MODES = [0, 1]  # MPM_SVD_EXACT, MPM_SVD_FAST


def _sim(N, mats, kind, mode=0, **kw):
    import mpm_b200

    return mpm_b200.Sim(N, DT, mats, model=kind, svd_mode=mode, **kw)


def _by_id(p_sorted_ids, arr):
    out = np.empty_like(arr)
    out[p_sorted_ids] = arr
    return out


@pytest.mark.parametrize("count", [0, 1, 31, 2049, 100_003])
def test_sort_bit_exact(count):
    """Stage (1): keys and permutation identical to a stable sort on the oracle's keys."""
    N = 32
    rng = np.random.default_rng(count)
    x = rng.uniform(-0.05, 1.05, (count, 3)).astype(np.float32)  # includes out-of-domain particles
    p = ol.new_particles(x)
    mats = ol.make_material(2e-6)
    sim = _sim(N, mats, ol.SNOW)
    sim.upload(p)
    keys, ids = sim.sort_state()
    ko = ol.cell_keys(p, DT, N)
    perm = ol.sort_perm(ko)
    assert np.array_equal(ids, perm)
    assert np.array_equal(keys, ko[perm])
    back = sim.download()
    assert back.tobytes() == p.tobytes()  # upload -> sort -> download restores upload order bit for bit


@pytest.mark.parametrize("speed,expect_merge", [(4.0, True), (60.0, False)])
def test_rebin_inside_a_substep_is_the_stable_sort(speed, expect_merge):
    """A due re-bin runs between the grid update and G2P with the cell keys written by that substep's
    P2G kernel (positions at the start of the substep).  Whether it merges (few particles changed cell:
    the moved ones are sorted and merged into the rest) or radix-sorts (many did), keys and permutation
    must be THE stable sort of the previous order by the new keys, bit for bit, and the state must still
    match the checker afterwards."""
    N = 32
    p, mats = scenes.two_spheres(N, kind=ol.SNOW, perturb=False)
    p["v"][:, 0] += speed
    sim = _sim(N, mats, ol.SNOW, 0, sort_every=3, graph_mode=1)   # (graphs off: a captured call cannot read the crossing count back)
    sim.upload(p)
    sim.advance(3)
    before = sim.download()          # positions the 4th substep starts with
    _, ids0 = sim.sort_state()       # the order the re-bin starts from
    r0, m0 = sim.rebins, sim.merge_rebins
    sim.advance(1)                   # re-bin due: happens inside this substep
    assert sim.rebins == r0 + 1
    assert (sim.merge_rebins == m0 + 1) == expect_merge
    keys, ids = sim.sort_state()
    ko = ol.cell_keys(before, DT, N)                  # new key of every particle, by id
    order = ol.sort_perm(ko[ids0])                    # stable sort of the previous order by the new keys
    assert np.array_equal(ids, ids0[order])
    assert np.array_equal(keys, ko[ids])
    changed = (ko != ol.cell_keys(p, DT, N)).mean()
    assert changed > 0.005, changed                   # the scene did change cells since the upload
    if expect_merge:
        ref, _ = ol.advance(p.copy(), mats, DT, N, ol.SNOW, 4)
        got = sim.download()
        assert np.abs(got["x"].astype(np.float64) - ref["x"]).max() * N < 1e-4
    # and again: this re-bin's output is the next one's "old keys" (the re-bin substep counts as the first of three)
    sim.advance(2)
    before = sim.download()
    _, ids0 = sim.sort_state()
    sim.advance(1)
    keys, ids = sim.sort_state()
    ko = ol.cell_keys(before, DT, N)
    assert np.array_equal(ids, ids0[ol.sort_perm(ko[ids0])]) and np.array_equal(keys, ko[ids])


def test_merge_rebin_random_displacements():
    """The merge re-bin on a dense block whose particles are displaced by hand between two re-bins (a few
    per cent of them by several cells in every direction, duplicates of keys, moves across tile
    boundaries), against the stable sort."""
    N, n = 64, 300_000
    p, mats = scenes.dense_block(n, N)
    sim = _sim(N, mats, ol.FIXED_COROTATED, 1, sort_every=2, graph_mode=1)
    sim.upload(p)
    sim.advance(1)
    rng = np.random.default_rng(3)
    cur = sim.download()
    pick = rng.random(n) < 0.04
    cur["x"][pick] += rng.uniform(-3.0 / N, 3.0 / N, (int(pick.sum()), 3)).astype(np.float32)
    cur["x"] = np.clip(cur["x"], 0.06, 0.94)
    sim.overwrite(cur)               # same slots, new positions: the order is stale now
    _, ids0 = sim.sort_state()
    m0 = sim.merge_rebins
    sim.advance(1)                   # not due yet (1 of 2)
    before = sim.download()
    sim.advance(1)                   # due: the keys come from `before`
    keys, ids = sim.sort_state()
    ko = ol.cell_keys(before, DT, N)
    assert np.array_equal(ids, ids0[ol.sort_perm(ko[ids0])])
    assert np.array_equal(keys, ko[ids])
    print("merge re-bins:", sim.merge_rebins - m0)


def test_dense_block_generator_matches_host():
    import mpm_b200

    n = 50_000
    p, mats = scenes.dense_block(n, 32)
    sim = _sim(32, mats, ol.FIXED_COROTATED)
    sim.generate_dense_block(n, seed=1234)
    got = sim.download()
    assert got.tobytes() == p.tobytes()


@pytest.mark.parametrize("kind", [ol.SNOW, ol.FIXED_COROTATED, ol.JELLY])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("p2g_mode", [0, 1])  # MPM_P2G_RUNS, MPM_P2G_DIRECT
@pytest.mark.parametrize("N", [32, 60])        # 60 = the reference's default grid: dx_inv = 59.9999962, not N
def test_p2g_single_step(kind, mode, p2g_mode, N):
    p, mats = scenes.two_spheres(N, kind=kind)
    sim = _sim(N, mats, kind, mode, p2g_mode=p2g_mode)
    sim.upload(p)
    sim.stage("reset_grid")
    sim.stage("p2g")
    g = sim.grid()
    go = ol.p2g(p, mats, DT, N, kind)
    # Atomics reorder sums: not bit-exact.  Per NODE and channel, against the sum of the magnitudes of
    # the terms that node received (w (|m v| + sum |A d|): m v and A d may cancel, rounding errors do
    # not) — SURVEY.md 8(c)(3): 1e-5 in exact mode.  Fast mode: 2.5e-4.  Its Newton polar rotation is
    # accurate to ~1e-7 while the reference's svd3 (4 Jacobi sweeps) leaves ~1e-6 on R; the stress
    # 2 mu (F - R) F^T divides that by the strain (2 % here), so stress-dominated nodes see the
    # REFERENCE's own svd3 error at ~1e-4 (measured 1.0e-4 .. 2.0e-4 at N = 60).  Plus one float epsilon of the busiest node of the channel: a node
    # that only sees the vanishing tail of a particle's spline (w ~ 1e-6) has its weight itself known
    # to ~1e-3 only in f32, whatever the evaluation order (x * dx_inv - base contracts to one FMA on the
    # GPU, also in the reference's own build; the checker rounds the product first).
    mag = ol.p2g_magnitudes(p, mats, DT, N, kind).astype(np.float64)
    tol = 1e-5 if mode == 0 else 2.5e-4
    err = np.abs(g.astype(np.float64) - go)
    bound = tol * mag + 1.2e-7 * mag.max(axis=(0, 1, 2), keepdims=True)
    worst = float((err / np.maximum(bound, 1e-300)).max())
    loaded = mag > 1e-3 * mag.max(axis=(0, 1, 2), keepdims=True)
    rel_loaded = float((err[loaded] / mag[loaded]).max())
    print(f"P2G N={N} kind={kind} mode={mode} p2g_mode={p2g_mode}: worst error / bound = {worst:.2f}, "
          f"worst relative error at nodes above 1e-3 of the busiest = {rel_loaded:.2e}")
    assert worst <= 1.0, worst
    assert ((g != 0) == (go != 0))[..., 3].all()   # the same nodes are touched
    # invariants (no checker): total mass of the particles whose whole stencil is inside the domain
    mass = float(mats[1])
    assert abs(g[..., 3].sum(dtype=np.float64) - mass * len(p)) <= 1e-6 * mass * len(p)


def test_grid_update_matches_oracle():
    N = 32
    p, mats = scenes.two_spheres(N)
    go = ol.p2g(p, mats, DT, N, ol.SNOW)
    sim = _sim(N, mats, ol.SNOW)
    sim.upload(p)
    sim.set_grid(go)
    sim.stage("grid_update")
    g = sim.grid()
    gu = ol.grid_update(go.copy(), DT, N)
    # velocities: identical division; gravity add contracts to one FMA on the GPU (<= 1 ulp)
    assert np.allclose(g[..., :3], gu[..., :3], rtol=3e-7, atol=1e-9)
    assert ((g[..., :3] == 0) == (gu[..., :3] == 0)).all()  # same wall planes / floor decisions
    # mass is left in place here (the reference overwrites it with 1.0)
    assert np.array_equal(g[..., 3], go[..., 3])


@pytest.mark.parametrize("kind", [ol.SNOW, ol.FIXED_COROTATED, ol.JELLY])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("g2p_mode", [0, 1])  # MPM_G2P_TILE, MPM_G2P_DIRECT
@pytest.mark.parametrize("N", [32, 60])
def test_g2p_single_step_from_identical_grid(kind, mode, g2p_mode, N):
    p, mats = scenes.two_spheres(N, kind=kind)
    go = ol.grid_update(ol.p2g(p, mats, DT, N, kind), DT, N)
    sim = _sim(N, mats, kind, mode, g2p_mode=g2p_mode)
    sim.upload(p)
    sim.set_grid(go)
    sim.stage("g2p")
    got = sim.download()
    ref = ol.g2p(go, p.copy(), mats, DT, N, kind)
    tol = 1e-5 if mode == 0 else 5e-5
    # N = 32: dx is a power of two, x * dx_inv and node * dx are exact and the only difference to the
    # checker is the order of the 27-term sums.  N = 60 (dx_inv = 59.9999962): nvcc contracts
    # x * dx_inv - base and node * dx - x into single FMAs (the reference's own build does too) while the
    # checker rounds the products first, which moves every weight and distance by ~1e-7 relative:
    # "the reference within FMA noise" (Oracle A caveat), three times the N = 32 bounds.
    fma_noise = 1.0 if N == 32 else 3.0
    assert scenes.rel_err(got["x"], ref["x"], 1e-2).max() < 1e-6
    assert scenes.rel_err(got["v"], ref["v"], 1e-2).max() < tol * fma_noise
    # C = 4/dx^2 * sum_i w v_i d_i^T: 27 terms of magnitude |v| * 4N/... that cancel; the summation
    # order differs (separable accumulation), so the bound is relative to the term magnitude
    c_scale = 4.0 * N * np.abs(go[..., :3]).max()
    assert np.abs(got["C"].astype(np.float64) - ref["C"]).max() < 1e-6 * c_scale * fma_noise
    assert scenes.rel_err(got["F"], ref["F"], 1.0).max() < tol * fma_noise
    assert scenes.rel_err(got["Jp"], ref["Jp"], 1.0).max() < tol * fma_noise


@pytest.mark.parametrize("kind,steps", [(ol.SNOW, 100), (ol.FIXED_COROTATED, 100)])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("sort_every,g2p_mode,pipeline", [(10, 0, 0), (0, 0, 0), (10, 0, 1), (0, 0, 1), (10, 1, 0)])
def test_short_horizon_against_oracle(kind, steps, mode, sort_every, g2p_mode, pipeline):
    """100 substeps of a two-ball impact: per-particle position / velocity error, mass, momentum.
    sort_every=0 never re-bins after the upload: the thrown ball drifts ~1 cell, so the kernels work
    on a stale order.  pipeline 0 = hand-over (default: G2P leaves the next P2G's affine matrix in the C
    rows between the substeps of one mpm_advance call), 1 = classic (every P2G evaluates the material);
    the hand-over run advances in calls of 1, 7 and the rest, so both forms of the particle state meet
    every kind of substep."""
    N = 32
    p, mats = scenes.two_spheres(N, kind=kind, perturb=False)
    sim = _sim(N, mats, kind, mode, sort_every=sort_every, g2p_mode=g2p_mode, pipeline=pipeline)
    sim.upload(p)
    sim.advance(1)
    sim.advance(7)
    sim.advance(steps - 8)
    got = sim.download()
    ref, _ = ol.advance(p.copy(), mats, DT, N, kind, steps)
    dx = 1.0 / N
    pos_err = np.abs(got["x"].astype(np.float64) - ref["x"]).max() / dx
    vel_err = np.linalg.norm(got["v"].astype(np.float64) - ref["v"], axis=1) / np.maximum(np.linalg.norm(ref["v"], axis=1), 1e-1)
    assert pos_err < 1e-3, pos_err          # |dx|/dx <= 1e-3
    assert np.quantile(vel_err, 0.999) < 1e-2, np.quantile(vel_err, 0.999)
    mom_g = got["v"].astype(np.float64).sum(0)
    mom_r = ref["v"].astype(np.float64).sum(0)
    assert np.abs(mom_g - mom_r).max() <= 1e-4 * np.abs(ref["v"]).astype(np.float64).sum()


def test_g2p_tile_handles_stale_order_and_domain_faces():
    """The staged G2P against the direct one from the same grid: particles scattered over the whole
    domain (clipped stencils at the faces, out-of-domain particles) and an order made stale by
    moving every particle up to 2.5 cells after the re-bin, so all of the window rows (di, dj in
    -1..1), the slack along z and the out-of-window path are exercised."""
    N = 32
    rng = np.random.default_rng(11)
    n = 40_000
    x0 = rng.uniform(-0.02, 1.02, (n, 3)).astype(np.float32)
    p = ol.new_particles(x0)
    p["F"] += 0.02 * rng.standard_normal(p["F"].shape).astype(np.float32)
    mats = ol.make_material(1.0 / 200000.0)
    g = rng.standard_normal((N, N, N, 4)).astype(np.float32)
    g[..., 3] = np.abs(g[..., 3])
    moved = p.copy()
    moved["x"] = x0 + rng.uniform(-2.5 / N, 2.5 / N, (n, 3)).astype(np.float32)
    out = []
    for g2p_mode in (0, 1):
        sim = _sim(N, mats, ol.SNOW, 0, g2p_mode=g2p_mode)
        sim.upload(p)          # bins by x0
        sim.overwrite(moved)   # same slots, new positions: the order and the key array are now stale
        sim.set_grid(g)
        sim.stage("g2p")
        out.append(sim.download())
    a, b = out
    for f in ("x", "v", "F", "C", "Jp"):
        assert np.allclose(a[f], b[f], rtol=2e-5, atol=2e-5 * np.abs(b[f]).max()), f
    ref = ol.g2p(g, moved.copy(), mats, DT, N, ol.SNOW)
    assert np.allclose(a["v"], ref["v"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("kind", [ol.SNOW, ol.FIXED_COROTATED])
@pytest.mark.parametrize("sort_every", [0, 3])
def test_handover_pipeline_matches_classic(kind, sort_every):
    """Hand-over (G2P computes the next P2G's affine matrix) against the classic pipeline (every P2G
    evaluates the material itself): particles over the whole domain (clipped stencils, out-of-domain
    particles that both paths must leave untouched, fast particles at a domain face), fast enough to change cells between re-bins, several materials,
    fixed-corotated particles with Jp != 1, and particle data replaced / single stages run between the
    calls.  Same arithmetic per particle, so the only difference is the order of the atomic sums."""
    N = 32
    rng = np.random.default_rng(23)
    p, _ = scenes.two_spheres(N, kind=kind)
    nb = int((p["x"][:, 1] > 0.45).sum())
    p["v"][p["x"][:, 1] > 0.45] = [5.0, -30.0, 2.0]   # the thrown ball crosses a cell in ~10 substeps
    p["C"] *= 0.2
    # a resting sheet that pokes through the x = 0 face (clipped stencils, sticky-wall nodes), a few
    # particles far outside the domain, which both pipelines must leave untouched, and a few just inside
    # the z = 1 face flying outwards (clipped stencils, sticky wall)
    sheet = ol.new_particles(rng.uniform([-0.045, 0.3, 0.3], [0.08, 0.5, 0.5], (4000, 3)).astype(np.float32))
    leaving = ol.new_particles(rng.uniform([0.3, 0.3, 1.0 - 0.3 / N], [0.5, 0.5, 1.0 - 0.1 / N], (200, 3)).astype(np.float32))
    leaving["v"][:, 2] = 60.0
    far = ol.new_particles(rng.uniform(1.05, 1.2, (64, 3)).astype(np.float32))
    far["v"] = 1.0
    p = np.concatenate([p, sheet, leaving, far])
    n = len(p)
    p["material_type"] = rng.integers(0, 2, n).astype(np.uint8)
    vol = 1.0 / 500000.0
    if kind == ol.SNOW:
        mats = np.stack([ol.make_material(vol), ol.make_material(vol, 400.0, 1.0e5, 0.3, 5.0, 0.97, 1.01)])
    else:
        p["Jp"] = (1.0 + 0.05 * rng.standard_normal(n)).astype(np.float32)  # the fixed-corotated lambda term
        mats = np.stack([ol.make_material(vol, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30),
                         ol.make_material(vol, 500.0, 0.7e5, 0.3, 0.0, 0.0, 1e30)])
    outs = []
    for pipeline in (0, 1):
        sim = _sim(N, mats, kind, 0, sort_every=sort_every, pipeline=pipeline)
        sim.upload(p)
        if kind != ol.SNOW:
            assert sim.diagnostics()["jp_not_one"] == 1
        sim.advance(4)
        mid = sim.download()
        mid["v"][:-64, 1] += np.float32(0.25)
        sim.overwrite(mid)        # particles changed between two calls
        sim.advance(3)
        sim.stage("reset_grid")   # single stages in between
        sim.stage("p2g")
        sim.advance(3)
        outs.append(sim.download())
        sim.close()
    a, b = outs
    assert a[-64:].tobytes() == p[-64:].tobytes() and b[-64:].tobytes() == p[-64:].tobytes()
    assert nb > 1000 and np.isfinite(a["x"]).all() and np.abs(a["x"][:-264]).max() < 1.0
    # (the sticky wall planes stop the fast particles at the z = 1 face: grid velocities there are zero,
    # so in practice nothing leaves the domain; the kernels still guard the hand-over against it)
    errs = {f: float(np.abs(a[f].astype(np.float64) - b[f]).max() / max(np.abs(b[f]).max(), 1e-6)) for f in ("x", "v", "F", "C", "Jp")}
    print("hand-over vs classic, max error / field maximum:", {f: f"{e:.1e}" for f, e in errs.items()})
    # Two runs of EITHER pipeline differ by the order of the atomic sums (~1e-7 per substep), which the
    # stiff two-ball contact amplifies over the 10 substeps; measured 1e-6 .. 2e-5 depending on the run.
    # A particle left with the affine matrix in place of C would show as O(1).
    for f, e in errs.items():
        assert e <= 2e-4, (f, errs)
    # against the oracle on the same schedule
    ref, _ = ol.advance(p.copy(), mats, DT, N, kind, 4)
    ref["v"][:-64, 1] += np.float32(0.25)
    ref, _ = ol.advance(ref, mats, DT, N, kind, 6)
    assert np.abs(a["x"].astype(np.float64) - ref["x"]).max() * N < 1e-3
    assert np.abs(a["v"].astype(np.float64) - ref["v"]).max() <= 1e-3 * np.abs(ref["v"]).max()
    c_scale = np.abs(ref["C"]).max()
    assert np.abs(a["C"].astype(np.float64) - ref["C"]).max() <= 2e-3 * c_scale


def test_adaptive_rebin_on_measured_disorder():
    """MpmParams.rebin_permille: re-bin when the cell crossings counted by G2P exceed a share of the
    particles.  A thrown ball triggers re-bins without any fixed cadence and the result still matches
    the oracle; a ball at rest in free fall for a few substeps never re-bins after the upload."""
    N, steps = 32, 100
    p, mats = scenes.two_spheres(N, kind=ol.SNOW, perturb=False)
    sim = _sim(N, mats, ol.SNOW, 0, sort_every=0, rebin_permille=20)
    sim.upload(p)
    assert sim.rebins == 1
    sim.advance(steps)
    got = sim.download()
    n_rebins = sim.rebins
    assert 2 <= n_rebins < steps // 2, n_rebins
    ref, _ = ol.advance(p.copy(), mats, DT, N, ol.SNOW, steps)
    assert np.abs(got["x"].astype(np.float64) - ref["x"]).max() * N < 1e-3
    rest = p[p["v"][:, 1] == 0].copy()
    sim2 = _sim(N, mats, ol.SNOW, 0, sort_every=0, rebin_permille=20)
    sim2.upload(rest)
    sim2.advance(10)
    sim2.sync()
    assert sim2.rebins == 1
    print(f"adaptive re-bin: {n_rebins - 1} re-bins in {steps} substeps of the two-ball impact")


def test_positions_readback_blocking_and_async():
    """positions-only read-back (12 B/particle) in upload order, blocking and queued."""
    N = 32
    p, mats = scenes.two_spheres(N)
    sim = _sim(N, mats, ol.SNOW, sort_every=2)
    sim.upload(p)
    sim.advance(3)
    out = np.zeros((len(p), 3), np.float32)
    assert sim.download_positions_async(out) == len(p)
    sim.advance(2)          # queued behind the copy: must not disturb it
    sim.sync()
    ref, _ = ol.advance(p.copy(), mats, DT, N, ol.SNOW, 3)
    assert np.abs(out.astype(np.float64) - ref["x"]).max() * N < 1e-4
    assert np.array_equal(sim.download_positions(), sim.download()["x"])


def test_overlapped_transfers_give_the_blocking_results():
    """mpm_prefetch_particles_aos / mpm_download_particles_aos_async: particle sets streamed through one handle
    with the copies of one set overlapping the substeps of another give what the blocking calls give."""
    import torch

    N, steps, n_sets = 32, 6, 4
    base, mats = scenes.two_spheres(N)
    n = len(base)
    sets = []
    for k in range(n_sets):
        q = base.copy()
        q["v"][:, 1] += 0.25 * k          # different inputs, so that a stale buffer would show
        sets.append(q)

    def pinned(src=None):
        t = torch.empty(n * 104, dtype=torch.uint8, pin_memory=True)
        a = t.numpy().view(ol.PARTICLE_DTYPE)
        if src is not None:
            a[:] = src
        return t, a

    want = []
    sim = _sim(N, mats, ol.SNOW, sort_every=2)
    for q in sets:
        sim.upload(q)
        sim.advance(steps)
        want.append(sim.download())
    sim.close()

    ins = [pinned(q) for q in sets]
    outs = [pinned() for _ in sets]
    sim = _sim(N, mats, ol.SNOW, sort_every=2)
    sim.prefetch_ptr(ins[0][0].data_ptr(), n)
    for k in range(n_sets):
        sim.upload_ptr(ins[k][0].data_ptr(), n)                   # consumes the prefetched copy
        if k + 1 < n_sets:
            sim.prefetch_ptr(ins[k + 1][0].data_ptr(), n)          # travels while set k is simulated
        sim.advance(steps)
        assert sim.download_ptr_async(outs[k][0].data_ptr(), n) == n   # travels while set k + 1 is simulated
    sim.download_wait()
    for k in range(n_sets):
        for f in ("x", "v", "F", "C", "Jp"):   # (not bit for bit: the order of P2G's float reductions differs from run to run)
            assert np.allclose(outs[k][1][f], want[k][f], rtol=0, atol=2e-4 if f == "C" else 2e-5), (k, f)
    # a prefetch that no upload takes up is harmless, as is an upload of another buffer in between
    sim.prefetch_ptr(ins[1][0].data_ptr(), n)
    sim.upload(sets[2])
    sim.advance(steps)
    got = sim.download()
    for f in ("x", "v", "F", "C", "Jp"):
        assert np.allclose(got[f], want[2][f], rtol=0, atol=2e-4 if f == "C" else 2e-5), f
    sim.close()


def test_remove_particles_is_a_stable_compaction():
    """mpm_remove_particles: the survivors keep their state and their cell order, the upload order closes up,
    and the run continues as if the removed object had never been uploaded (it is far from the others here)."""
    N = 32
    rng = np.random.default_rng(9)
    a = ol.new_particles(ol.sphere_positions(200000.0, 0.12, [0.3, 0.5, 0.3], rng))
    b = ol.new_particles(ol.sphere_positions(200000.0, 0.10, [0.7, 0.5, 0.7], rng))   # removed later
    c = ol.new_particles(ol.sphere_positions(200000.0, 0.08, [0.3, 0.3, 0.7], rng))
    for q, vy in ((a, 1.0), (b, -2.0), (c, 0.5)):
        q["v"][:, 1] = vy
    mats = ol.make_material(1.0 / 200000.0)
    everything = np.concatenate([a, b, c])
    sim = _sim(N, mats, ol.SNOW, sort_every=4)
    sim.upload(everything)
    sim.advance(6)
    before = sim.download()
    k0, ids0 = sim.sort_state()
    sim.remove(len(a), len(b))
    assert sim.count == len(a) + len(c)
    after = sim.download()
    keep = np.r_[0:len(a), len(a) + len(b):len(everything)]
    for f in ("x", "v", "F", "C", "Jp"):
        assert np.array_equal(after[f], before[f][keep]), f
    _, ids1 = sim.sort_state()
    alive = (ids0 < len(a)) | (ids0 >= len(a) + len(b))
    want_ids = ids0[alive]
    want_ids = np.where(want_ids >= len(a) + len(b), want_ids - len(b), want_ids)
    assert np.array_equal(ids1, want_ids)                        # same relative (cell) order, ids closed up
    sim.advance(10)                                              # crosses re-bins (a full sort: the old keys are gone)
    got = sim.download()
    ref, _ = ol.advance(before[keep].copy(), mats, DT, N, ol.SNOW, 10)
    assert np.abs(got["x"].astype(np.float64) - ref["x"]).max() * N < 1e-4
    assert np.abs(got["v"].astype(np.float64) - ref["v"]).max() < 1e-3
    sim.remove(0, sim.count)                                     # everything: an empty handle keeps working
    assert sim.count == 0
    sim.advance(2)
    sim.upload(c)
    sim.advance(2)
    assert sim.count == len(c)
    sim.close()


def test_free_fall_velocity():
    """Oracle-free invariant: before contact v_y(t) = -9.81 t (SURVEY.md 8(c) pin 6)."""
    N = 32
    rng = np.random.default_rng(5)
    x = ol.sphere_positions(200000.0, 0.2, [0.4, 0.5, 0.4], rng)
    p = ol.new_particles(x)
    mats = ol.make_material(1.0 / 200000.0)
    sim = _sim(N, mats, ol.SNOW)
    sim.upload(p)
    sim.advance(50)
    got = sim.download()
    assert np.allclose(got["v"][:, 1], -9.81 * 50 * DT, rtol=2e-3)
    assert np.abs(got["v"][:, [0, 2]]).max() < 1e-3


def test_cuda_graph_replay_of_advance_calls():
    """MpmParams.graph_mode: the launches of mpm_advance(n) are captured once per state of the re-bin
    cadence and replayed; the result is the ordinary one (against the checker), interleaved API calls
    drop the graphs, and the replays are counted."""
    import mpm_b200

    N, steps = 32, 192
    p, mats = scenes.two_spheres(N, kind=ol.SNOW, perturb=False)
    ref, _ = ol.advance(p.copy(), mats, DT, N, ol.SNOW, steps)
    out = {}
    for mode in (mpm_b200.GRAPH_ON, mpm_b200.GRAPH_OFF):
        sim = _sim(N, mats, ol.SNOW, 0, sort_every=8, graph_mode=mode)
        sim.upload(p)
        for _ in range(steps // 12):   # 12 substeps per call against a cadence of 8: a handful of (phase, buffer parity) states recur
            sim.advance(12)
        out[mode] = (sim.download(), sim.graph_replays, sim.rebins, sim.launches)
        sim.close()
    on, off = out[mpm_b200.GRAPH_ON], out[mpm_b200.GRAPH_OFF]
    assert on[1] >= steps // 24 and off[1] == 0                # every state seen before is a replay
    assert on[2] == off[2]                                     # same re-bins (the launch counts differ: a capture re-bins by radix sort)
    for got in (on[0], off[0]):
        assert np.abs(got["x"].astype(np.float64) - ref["x"]).max() * N < 3e-3   # 192 substeps of the impact
    # an upload in between drops the graphs; the handle keeps working
    sim = _sim(N, mats, ol.SNOW, 0, sort_every=8, graph_mode=mpm_b200.GRAPH_ON)
    sim.upload(p)
    sim.advance(8)
    sim.advance(8)
    sim.upload(p)
    sim.advance(8)
    got = sim.download()
    ref8, _ = ol.advance(p.copy(), mats, DT, N, ol.SNOW, 8)
    assert np.abs(got["x"].astype(np.float64) - ref8["x"]).max() * N < 1e-4
