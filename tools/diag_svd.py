"""Prints the deviation of the FAST svd3 policy from the EXACT one (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import mpm_b200, oracle_lib as ol

rng = np.random.default_rng(3)
n = 200_000
A = (np.eye(3, dtype=np.float32).reshape(9) + 0.05 * rng.standard_normal((n, 9))).astype(np.float32)
for mode in (0, 1):
    U, S, V = mpm_b200.svd3_batch(A, mode)
    Uo, So, Vo = ol.svd3(A)
    rec = np.einsum("nij,nj,nkj->nik", U.astype(np.float64), S.astype(np.float64), V.astype(np.float64))
    err = np.abs(rec - A.reshape(-1, 3, 3)).sum((1, 2))
    R, Ro = np.einsum("nij,nkj->nik", U, V), np.einsum("nij,nkj->nik", Uo, Vo)
    orth = np.abs(np.einsum("nij,nkj->nik", R.astype(np.float64), R.astype(np.float64)) - np.eye(3)).sum((1, 2))
    print(f"mode {mode}: recon L1 max {err.max():.3e} mean {err.mean():.3e} | dS max {np.abs(S-So).max():.3e} | dR max {np.abs(R-Ro).max():.3e} mean {np.abs(R-Ro).mean():.3e} | RRt-I L1 max {orth.max():.3e}")
M1 = np.array([[1, 2, 1], [1, 3, 1], [1, 8, 1]], np.float32)
M2 = np.array([[0, 1, 0], [-1, 2, -1], [-1, 0.001, -1]], np.float32)
for M in (M1, M2):
    for mode in (0, 1):
        U, S, V = mpm_b200.svd3_batch(M, mode)
        R = mpm_b200.polar_batch(M, mode)[0]
        Sym = (V[0] * S[0]) @ V[0].T
        print("gtest", mode, np.abs(R @ Sym - M).sum(), np.abs(R @ R.T - np.eye(3)).sum(), np.abs(Sym - Sym.T).sum(), S[0])
