"""Small deterministic scenes shared by the CPU and GPU parity tests."""
import numpy as np

import oracle_lib as ol


def two_spheres(N=32, density=500000.0, seed=1, kind=ol.SNOW, perturb=True):
    """A resting ball and a smaller ball thrown at it (snowman-like, sphere.obj stand-in)."""
    rng = np.random.default_rng(seed)
    xa = ol.sphere_positions(density, 0.4, [0.3, 0.12, 0.3], rng)
    xb = ol.sphere_positions(density, 0.2, [0.4, 0.6, 0.4], rng)
    p = ol.new_particles(np.concatenate([xa, xb]))
    p["v"][len(xa):] = [0.5, -3.0, 0.2]
    if perturb:  # exercise F, C, Jp paths from step one
        p["F"] += 0.02 * rng.standard_normal(p["F"].shape).astype(np.float32)
        p["C"] = 5.0 * rng.standard_normal(p["C"].shape).astype(np.float32)
        p["Jp"] = 1.0 + 0.05 * rng.standard_normal(len(p)).astype(np.float32)
    if kind == ol.SNOW:
        mats = ol.make_material(1.0 / density)  # scenes/snowman.toml
    elif kind == ol.JELLY:
        mats = ol.make_material(1.0 / density, 1000.0, 1.0e5, 0.3, 10.0, 0.0, 1e30)  # MMJelly's defaults
    else:  # fixed-corotated keeps the perturbed Jp: its lambda term reads it (MaterialModel.cuh:59-60)
        mats = ol.make_material(1.0 / density, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
    return p, mats


def dense_block(count, N, density=None, seed=1234, kind=ol.FIXED_COROTATED, shear=0.0, f_noise=0.0, first_id=0, lo=0.1, hi=0.9):
    """SURVEY.md 8(d) configs 4 / 5 at any size: uniform block in [0.1,0.9]^3, fixed-corotated or the
    snowman's snow; optionally under stress (mpm_generate_dense_block_stressed)."""
    x = ol.dense_block_positions(count, seed, lo, hi, first_id=first_id)
    p = ol.new_particles(x)
    ol.dense_block_stress(p, seed, shear, f_noise, first_id)
    dens = density if density is not None else count / (hi - lo) ** 3
    if kind == ol.SNOW:
        mats = ol.make_material(1.0 / dens, 700.0, 1.4e5, 0.2, 10.0, 0.975, 1.0075)  # scenes/snowman.toml
    else:
        mats = ol.make_material(1.0 / dens, 1000.0, 1.4e5, 0.2, 0.0, 0.0, 1e30)
    return p, mats


def rel_err(a, b, floor):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
