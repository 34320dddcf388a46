"""Runs in its own process with MPM_B200_LIB = the library that has tests/plugin/user_material.cu
compiled in: the user-defined material (model ids 16 = staged kernels, 17 = generic kernels through a
user-owned interpolation kernel / transfer scheme) against the CPU checker's MMJelly.  Prints JSON."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import mpm_b200  # noqa: E402
import oracle_lib as ol  # noqa: E402
import scenes  # noqa: E402

N, DT, STEPS = 32, 1e-4, 40
p, mats = scenes.two_spheres(N, kind=ol.JELLY)
p["Jp"][:100] = 0.5     # outside the clamp of the end-of-step hook (0.6 .. 20): the hook must pull them in
p["Jp"][100:200] = 25.0
raw = np.ascontiguousarray(mats[:5], np.float32).tobytes()  # UserHardeningSolid: volume, mass, mu0, lambda0, hardening
out = {"staged_flag": {}}
for model in (16, 17):
    r = {}
    # P2G against the checker
    sim = mpm_b200.Sim(N, DT, None, model=model, svd_mode=mpm_b200.SVD_EXACT, raw_materials=(raw, 1))
    sim.upload(p)
    sim.stage("reset_grid")
    sim.stage("p2g")
    g = sim.grid()
    go = ol.p2g(p, mats, DT, N, ol.JELLY)
    r["p2g"] = float(max(np.abs(g[..., c] - go[..., c]).max() / np.abs(go[..., c]).max() for c in range(4)))
    # G2P from the checker's grid
    gu = ol.grid_update(go.copy(), DT, N)
    sim.set_grid(gu)
    sim.stage("g2p")
    got = sim.download()
    ref = ol.g2p(gu, p.copy(), mats, DT, N, ol.JELLY)
    r["g2p_F"] = float(np.abs(got["F"].astype(np.float64) - ref["F"]).max())
    r["g2p_Jp"] = float(np.abs(got["Jp"].astype(np.float64) - ref["Jp"]).max())
    sim.close()
    # short horizon, hand-over pipeline where the model has the staged kernels
    sim = mpm_b200.Sim(N, DT, None, model=model, svd_mode=mpm_b200.SVD_EXACT, sort_every=7, raw_materials=(raw, 1))
    sim.upload(p)
    sim.advance(STEPS)
    got = sim.download()
    ref, _ = ol.advance(p.copy(), mats, DT, N, ol.JELLY, STEPS)
    r["x"] = float(np.abs(got["x"].astype(np.float64) - ref["x"]).max() * N)
    r["v"] = float(np.abs(got["v"].astype(np.float64) - ref["v"]).max())
    r["Jp_changed"] = float(np.abs(ref["Jp"] - p["Jp"]).max())
    r["launches"] = int(sim.launches)
    sim.close()
    out[str(model)] = r
# an unregistered id must be refused, and so must materials of the wrong size
try:
    mpm_b200.Sim(N, DT, None, model=18, raw_materials=(raw, 1))
    out["unregistered"] = "accepted"
except mpm_b200.MpmError as e:
    out["unregistered"] = str(e)
try:
    mpm_b200.Sim(N, DT, None, model=16, raw_materials=(raw + b"\0\0\0\0", 1))
    out["wrong_size"] = "accepted"
except mpm_b200.MpmError as e:
    out["wrong_size"] = str(e)
print(json.dumps(out))
