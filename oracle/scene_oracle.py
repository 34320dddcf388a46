"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the reference's scene front end, used only by
tests/ to check mpm_b200/host/*.hpp.  Never imported by the product.

Follows: src/main.cu:31-66 (TOML -> materials / objects; parsed here with Python's tomllib, an
independent parser), Simulation::loadMesh src/mpm.cu:331-346, Simulation::addParticles
src/mpm.cu:348-394 (glibc rand() through ctypes, x, y, z per point, 2048 points per batch, whole
batches drawn), igl::solid_angle include/igl/solid_angle.cpp:12-55 in float32, summed over the
faces in order (libigl sums through an AABB hierarchy: parity with the reference itself is
unpinned at points whose winding number is within float noise of 1).
"""
import ctypes
import math
import tomllib

import numpy as np

_libc = ctypes.CDLL("libc.so.6")
RAND_MAX = 2147483647
f32 = np.float32


def load_scene_toml(path):
    with open(path, "rb") as f:
        return tomllib.load(f)


def rescale(V, size, position):
    """loadMesh: float32 arithmetic on the vertex matrix, `size` a double."""
    V = np.asarray(V, f32).copy()
    mn, mx = V.min(0), V.max(0)
    length_max = (mx - mn).max()
    scale = f32(float(size) / float(length_max))
    shift = np.asarray(position, f32) - scale * mn
    return (V * scale + shift).astype(f32)


def solid_angles_2pi(A, B, C, P):
    """One triangle (A, B, C: float32[3]) seen from points P (n x 3): float32 like the reference."""
    v0, v1, v2 = (A - P).astype(f32), (B - P).astype(f32), (C - P).astype(f32)

    def norm(v):
        return np.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]).astype(f32) + v[:, 2] * v[:, 2]).astype(f32)

    vl0, vl1, vl2 = norm(v0), norm(v1), norm(v2)
    det = (v0[:, 0] * v1[:, 1] * v2[:, 2] + v1[:, 0] * v2[:, 1] * v0[:, 2] + v2[:, 0] * v0[:, 1] * v1[:, 2]
           - v2[:, 0] * v1[:, 1] * v0[:, 2] - v1[:, 0] * v0[:, 1] * v2[:, 2] - v0[:, 0] * v2[:, 1] * v1[:, 2]).astype(f32)

    def dot(a, b):
        d = (a[:, 0] * b[:, 0]).astype(f32)
        d = (d + a[:, 1] * b[:, 1]).astype(f32)
        return (d + a[:, 2] * b[:, 2]).astype(f32)

    dp0, dp1, dp2 = dot(v1, v2), dot(v2, v0), dot(v0, v1)
    den = (vl0 * vl1 * vl2 + dp0 * vl0 + dp1 * vl1 + dp2 * vl2).astype(f32)
    return (np.arctan2(det.astype(np.float64), den.astype(np.float64)) / (2.0 * math.pi)).astype(f32)


def winding_numbers(V, F, P):
    P = np.asarray(P, f32)
    w = np.zeros(len(P), f32)
    for a, b, c in F:
        w = (w + solid_angles_2pi(V[a], V[b], V[c], P)).astype(f32)
    return w


def add_particles(V, F, particle_density, rand=None):
    """Positions accepted by addParticles, in order.  rand: callable returning the next glibc rand()."""
    rand = rand or _libc.rand
    Vd = V.astype(np.float64)
    mn, mx = Vd.min(0), Vd.max(0)
    rng = mx - mn
    target = int(np.uint32(particle_density * (rng[0] * rng[1] * rng[2])))
    out, count = [], 0
    while count < target:
        r = np.array([rand() for _ in range(3 * 2048)], np.float64).reshape(2048, 3)
        u = (r.astype(f32) / f32(RAND_MAX)).astype(f32)  # float(rand()) / float(RAND_MAX)
        pts = (mn + u.astype(np.float64) * rng).astype(f32)
        W = winding_numbers(V, F, pts).astype(np.int32)  # float -> int truncation
        take = min(2048, target - count)
        out.append(pts[:take][W[:take] == 1])
        count += take
    return np.concatenate(out) if out else np.zeros((0, 3), f32)


def srand(seed):
    _libc.srand(ctypes.c_uint(seed))
