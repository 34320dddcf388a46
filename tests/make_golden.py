"""Generates tests/golden/*.npz from the reference's OWN code built for the host (oracle/_ref,
which exists only where /root/reference is mounted).  Run: python tests/make_golden.py
The fixtures let the GPU box (no /root/reference) check the oracle and the CUDA path against
outputs of the reference itself."""
import os

import numpy as np

import oracle_lib as ol
import scenes

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def svd3_golden():
    rng = np.random.default_rng(2024)
    A = rng.standard_normal((4096, 9)).astype(np.float32)
    A[:1024] = np.eye(3, dtype=np.float32).reshape(9) + 0.05 * A[:1024]
    A[1024:2048] *= np.float32(10) ** rng.uniform(-6, 6, (1024, 1)).astype(np.float32)
    A[-16:, 2] = A[-16:, 0]; A[-16:, 5] = A[-16:, 3]; A[-16:, 8] = A[-16:, 6]
    A[-32:-16] = 0
    A[-48:-32] = np.diag([1, 1, -0.9]).astype(np.float32).reshape(9)
    A[0] = [1, 2, 1, 1, 3, 1, 1, 8, 1]          # tests/test_linalg.cu:27
    A[1] = [0, 1, 0, -1, 2, -1, -1, 0.001, -1]  # tests/test_linalg.cu:33
    U, S, V = ol.svd3(A, which="ref")
    np.savez_compressed(os.path.join(OUT, "svd3_golden.npz"), A=A, U=U, S=S, V=V)


def substep_golden():
    for kind, name in ((ol.SNOW, "snow"), (ol.FIXED_COROTATED, "fc"), (ol.JELLY, "jelly")):
        ref = ol.Ref(kind)
        assert ref.available, "needs oracle/_ref (build where /root/reference is mounted)"
        N, dt = 16, 1e-4
        p, mats = scenes.two_spheres(N, density=40000.0, seed=7, kind=kind)
        assert len(p) <= N ** 3
        p0 = p.copy()
        g = ref.p2g(p, mats, dt, N)
        g_p2g = g.copy()
        ref.grid_update(g, dt, N)
        g_upd = g.copy()
        ref.g2p(g, p, mats, dt, N)
        p1 = p.copy()
        ref.advance(p, mats, dt, N, 20)
        np.savez_compressed(os.path.join(OUT, f"substep_{name}_golden.npz"), N=N, dt=dt, mats=mats, p0=p0,
                            grid_after_p2g=g_p2g, grid_after_update=g_upd, p1=p1, p21=p)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ol.set_threads(1)
    svd3_golden()
    substep_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
