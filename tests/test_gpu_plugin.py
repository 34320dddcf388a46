"""The plugin surface (include/mpm_b200/plugin.cuh): a user-defined material, and a user-owned
interpolation kernel / transfer scheme, compiled into the library and checked against the CPU
checker.  The build (mpm_b200/build.py: build_plugin_example) happens in __graft_entry__.build()."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_user_defined_material_against_checker():
    from mpm_b200 import build as b

    lib = b.build_plugin_example()
    env = dict(os.environ, MPM_B200_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "plugin", "run_user_material.py")], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    print(out)
    assert "no material model registered" in out["unregistered"]
    assert "bytes" in out["wrong_size"]
    for model in ("16", "17"):
        e = out[model]
        assert e["p2g"] < 1e-5, (model, e)        # atomics reorder sums
        assert e["g2p_F"] < 1e-5 and e["g2p_Jp"] < 1e-5, (model, e)
        assert e["x"] < 1e-3 and e["v"] < 5e-2, (model, e)
        assert e["Jp_changed"] > 0.09               # the end-of-step hook clamped the out-of-range Jp, in the checker and (g2p_Jp) here


def test_default_library_has_no_user_models():
    import mpm_b200

    with pytest.raises(mpm_b200.MpmError, match="no material model registered"):
        mpm_b200.Sim(32, 1e-4, mpm_b200.make_material(2e-6), model=16)
