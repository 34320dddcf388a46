"""Small end-to-end run that touches every kernel of the library, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/mini.py        (also racecheck / initcheck / synccheck)
Every model in both arithmetic policies, the hand-over and the classic pipeline, the staged and the generic
kernels, radix and merge re-bins, the stale-order P2G variant, graph replay, the on-device generator."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mpm_b200

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
P = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6


def run(model, svd, pipeline, p2g_mode, g2p_mode, graph_mode, shear):
    mats = mpm_b200.make_material(0.512 / P, 1000.0, 1.4e5, 0.2, 10.0, 0.975, 1.0075)
    sim = mpm_b200.Sim(N, 1e-4, mats, model=model, svd_mode=svd, sort_every=2, capacity=2 * P, pipeline=pipeline,
                       p2g_mode=p2g_mode, g2p_mode=g2p_mode, graph_mode=graph_mode)
    sim.generate_dense_block(P, seed=7, shear=shear, f_noise=0.03 if shear else 0.0)
    for _ in range(3):
        sim.advance(steps)
    got = sim.download()
    d = sim.diagnostics()
    assert np.isfinite(got["x"]).all() and d["nonfinite"] == 0, d
    out = (sim.launches, sim.rebins, sim.merge_rebins, sim.graph_replays)
    sim.close()
    return out


for model in (mpm_b200.FIXED_COROTATED, mpm_b200.SNOW, mpm_b200.JELLY):
    for svd in (mpm_b200.SVD_EXACT, mpm_b200.SVD_FAST):
        for pipeline in (mpm_b200.PIPE_HANDOVER, mpm_b200.PIPE_CLASSIC):
            print(model, svd, pipeline, run(model, svd, pipeline, mpm_b200.P2G_RUNS, mpm_b200.G2P_TILE, mpm_b200.GRAPH_OFF, 0.0))
print("sheared", run(mpm_b200.SNOW, mpm_b200.SVD_FAST, mpm_b200.PIPE_HANDOVER, mpm_b200.P2G_RUNS, mpm_b200.G2P_TILE, mpm_b200.GRAPH_OFF, 200.0))
print("graphs", run(mpm_b200.SNOW, mpm_b200.SVD_FAST, mpm_b200.PIPE_HANDOVER, mpm_b200.P2G_RUNS, mpm_b200.G2P_TILE, mpm_b200.GRAPH_ON, 0.0))
print("generic", run(mpm_b200.SNOW, mpm_b200.SVD_EXACT, mpm_b200.PIPE_CLASSIC, mpm_b200.P2G_DIRECT, mpm_b200.G2P_DIRECT, mpm_b200.GRAPH_OFF, 0.0))


def transfers():
    """overlapped transfers, removal of an id range, positions read-back"""
    import torch
    mats = mpm_b200.make_material(0.512 / P, 1000.0, 1.4e5, 0.2, 10.0, 0.975, 1.0075)
    sim = mpm_b200.Sim(N, 1e-4, mats, model=mpm_b200.SNOW, svd_mode=mpm_b200.SVD_FAST, sort_every=2, capacity=2 * P)
    sim.generate_dense_block(P, seed=3)
    sim.advance(2)
    n = sim.count
    a = torch.empty(n * 104, dtype=torch.uint8, pin_memory=True)
    b = torch.empty(n * 104, dtype=torch.uint8, pin_memory=True)
    sim.download_ptr(a.data_ptr(), n)
    sim.prefetch_ptr(a.data_ptr(), n)
    for _ in range(3):
        sim.upload_ptr(a.data_ptr(), n)
        sim.prefetch_ptr(a.data_ptr(), n)
        sim.advance(steps)
        sim.download_ptr_async(b.data_ptr(), n)
    sim.download_wait()
    sim.remove(n // 3, n // 4)
    sim.advance(steps)
    xyz = sim.download_positions()
    assert np.isfinite(xyz).all() and len(xyz) == n - n // 4
    out = (sim.launches, sim.rebins, sim.count)
    sim.close()
    return out


print("transfers", transfers())
print("mini ok")
