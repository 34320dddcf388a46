"""Tiny end-to-end run for compute-sanitizer: python tools/mini.py [N] [particles] [steps] [fuse]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpm_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
P = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
fuse = int(sys.argv[4]) if len(sys.argv) > 4 else 0
for model in (mpm_b200.FIXED_COROTATED, mpm_b200.SNOW):
    mats = mpm_b200.make_material(0.512 / P, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
    sim = mpm_b200.Sim(N, 1e-4, mats, model=model, svd_mode=mpm_b200.SVD_FAST, sort_every=2, fuse_mode=fuse, capacity=2 * P)
    sim.generate_dense_block(P, seed=1)
    sim.advance(steps)
    d = sim.download()
    sim.append(np.ascontiguousarray(d[: P // 4]))   # an object entering
    sim.advance(steps)
    x = sim.download_positions()
    sim.sync()
    d = sim.download()
    print("ok", model, fuse, len(d), float(d["v"][:, 1].mean()), float(x.mean()))
    sim.close()
