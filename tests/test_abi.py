"""The C-ABI library loads and exports every symbol include/mpm_b200.h declares (no compute)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mpm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    import mpm_b200

    L = mpm_b200.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert L.mpm_abi_version() == 6


def test_layouts_match_reference_sizes():
    import mpm_b200
    from mpm_b200 import api

    assert mpm_b200.PARTICLE_DTYPE.itemsize == 104  # sizeof(MLS_APIC_Particle)
    assert api.MATERIAL_DTYPE.itemsize == 28        # sizeof(MMSnow<Particle>)
    off = {n: mpm_b200.PARTICLE_DTYPE.fields[n][1] for n in ("x", "v", "F", "C", "Jp")}
    assert off == {"x": 4, "v": 16, "F": 28, "C": 64, "Jp": 100}


def test_make_material_matches_oracle():
    import mpm_b200
    import oracle_lib as ol

    for args in ((2e-6,), (1e-5, 200.0, 1.4e5, 0.2, 0.0, 0.0, 1e30), (1e-6, 1000.0, 1.4e5, 0.45, 0.0, 0.975, 0.975)):
        assert np.array_equal(mpm_b200.make_material(*args), ol.make_material(*args))


def test_no_cpu_fallback():
    """Without a GPU the product must fail loudly, not compute on the host."""
    import torch

    import mpm_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mpm_b200.MpmError, match="no CUDA device"):
        mpm_b200.Sim(32, 1e-4, mpm_b200.make_material(2e-6))
    with pytest.raises(mpm_b200.MpmError):
        mpm_b200.svd3_batch(np.eye(3, dtype=np.float32))


def test_product_never_touches_the_checker():
    """Nothing under mpm_b200/ may import, link or execute the parity checker."""
    word = "ora" + "cle"
    for root, _, files in os.walk(os.path.join(ROOT, "mpm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert word not in open(os.path.join(root, f)).read().lower(), f
